# cfg-2 bench A/B over an environment switch given as $1 (stage table)
for v in "" "$1" "" "$1"; do
  echo "--- env: $v"
  env $v timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-reference-semantics 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][0]
d=json.loads(l)
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['gate_inlier_frac_last_step'])
print({k: round(v,4) for k,v in d['stage_ms_per_update'].items()})"
done
