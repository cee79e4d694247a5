timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^ok" | tail -30
( time timeout 600 python bench.py 2>gpurun_out/s2_bench2.err ) 2>&1 | tee gpurun_out/s2_bench2.json | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][0]
d=json.loads(l)
print(d['value'], d['ms_per_step'], 'e2e', d['e2e'], d['gate_inlier_frac_last_step'], d['gpu_launches'])
print(d['stage_ms_per_update']); print(json.dumps(d['roofline'], indent=1)); print(d['clocks']); print(d['cpu_baseline'])"
tail -3 gpurun_out/s2_bench2.json
