# one launch list + full captures of the top kernels of a steady-state cfg-2 update, and of the TMA-staged contraction
# kernels at cfg-5 dimensions (outputs under gpurun_out/)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-semantics"
T=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/${T}_launches.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tracks -s 36 -c 1 -o gpurun_out/${T}_tracks $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tallchol -s 75 -c 3 -o gpurun_out/${T}_tallchol $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_downdate_mma -s 60 -c 3 -o gpurun_out/${T}_downdate_mma $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm_mma -s 250 -c 7 -o gpurun_out/${T}_gemm_mma $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_prop_step -s 100 -c 1 -o gpurun_out/${T}_prop_step $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_omega_small -s 40 -c 1 -o gpurun_out/${T}_omega_small $B > /dev/null 2>&1
# cfg-5: the downdate (k_gemm_tma<true>) and the Schur complement (k_gemm_tma<false>) of the last full-size updates
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tma -s 57 -c 4 -o gpurun_out/${T}_gemm_tma python tools/cfg5_check.py > /dev/null 2>&1
ls -la gpurun_out/ | grep ${T}
