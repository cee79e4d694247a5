# ncu --set full of the TMA-staged contraction kernels at cfg-5 dimensions (downdate = k_gemm_tma<true>, Schur = <false>)
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tma -s 57 -c 4 -o gpurun_out/r02_gemm_tma python tools/cfg5_check.py > gpurun_out/tma_prof.log 2>&1
tail -3 gpurun_out/tma_prof.log
ls -la gpurun_out | grep gemm_tma
