# TMA-staged contraction kernels: unit tests, micro-benchmark, cfg-5 check (XB_NO_TMA=1 = cp.async kernels for A/B), ncu capture
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_vs_reference.py -q -x 2>&1 | tail -4
timeout 120 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.txt
timeout 300 python tools/cfg5_check.py 2>&1 | tail -3 | tee gpurun_out/cfg5_tma.txt
bash tools/gpu/tma_prof.sh
