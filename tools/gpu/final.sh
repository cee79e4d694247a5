python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py 2>gpurun_out/final_bench.err | tee gpurun_out/final_bench.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/final_ref.json | cut -c1-300
timeout 200 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.txt
timeout 300 python tools/cfg5_check.py 2>&1 | tail -4 | tee gpurun_out/cfg5_check.txt
bash tools/gpu/prof_all.sh r02 > /dev/null 2>&1
ls gpurun_out | head -40
