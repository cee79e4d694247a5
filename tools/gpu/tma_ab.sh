# TMA-staged contraction kernels: unit tests, micro-benchmark, cfg-5 A/B (XB_NO_TMA=1 = cp.async kernels), launch list at cfg-5
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -15
timeout 120 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.txt
echo "--- cfg-5, cp.async kernels"; XB_NO_TMA=1 timeout 300 python tools/cfg5_check.py 2>&1 | tail -4 | tee gpurun_out/cfg5_no_tma.txt
echo "--- cfg-5, TMA kernels"; timeout 300 python tools/cfg5_check.py 2>&1 | tail -4 | tee gpurun_out/cfg5_tma.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/cfg5_launches.csv python tools/cfg5_check.py > /dev/null 2>&1
tail -n 400 gpurun_out/cfg5_launches.csv > gpurun_out/cfg5_launches_tail.csv; rm -f gpurun_out/cfg5_launches.csv
