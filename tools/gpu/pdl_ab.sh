# programmatic dependent launch A/B: GPU tests, then the cfg-2 bench with and without the launch attribute
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for v in 0 1 0 1; do
  echo "--- XB_NO_PDL=$v"
  XB_NO_PDL=$v timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-reference-semantics 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][0]
d=json.loads(l)
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['gate_inlier_frac_last_step'])
print({k: round(v,4) for k,v in d['stage_ms_per_update'].items()})
print('cfg5', d.get('cfg5',{}).get('ms_per_update'))"
done
