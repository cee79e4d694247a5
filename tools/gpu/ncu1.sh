B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_downdate32 -s 45 -c 2 -o gpurun_out/s2_dd32 $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm32 -s 140 -c 6 -o gpurun_out/s2_gemm32 $B > /dev/null 2>&1
ls -la gpurun_out/
