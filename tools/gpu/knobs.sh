for cfg in "XB_NONE=1" "XB_CHOL_SHARE=5" "XB_CHOL_SHARE=6" "XB_CHOL_SHARE=8" "XB_NONE=2" "XB_CHOL_SHARE=6"; do
echo "== $cfg"
env $cfg timeout 300 python bench.py --no-cpu-baseline --no-reference-semantics --steps 60 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][0]
d=json.loads(l)
s=d['stage_ms_per_update']
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {k:s[k] for k in ('tracks','side_tallchol_slam_cols','side_slam_part','chol_gram','tallchol','gram')})"
done
