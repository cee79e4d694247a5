# round-end check of HEAD on one GPU: smoke, the GPU test suite, the default bench line and the reference arm
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py 2>gpurun_out/final_bench.err | tee gpurun_out/final_bench.json | cut -c1-300
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/final_ref.json | cut -c1-200
