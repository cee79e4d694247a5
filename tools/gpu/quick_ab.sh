# GPU tests + cfg-2 bench (stage table, IMU timings)
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 400 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-reference-semantics 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][0]
d=json.loads(l)
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', d['e2e'], d['gate_inlier_frac_last_step'])
print({k: round(v,4) for k,v in d['stage_ms_per_update'].items()})
print(d['imu_us_per_sample']['value'], d['imu_us_per_sample']['batched'])"
