for cfg in "XB_TRACK_WARPS=1" "XB_TRACK_WARPS=2" "XB_TRACK_WARPS=4" "XB_CHOL_SHARE=2" "XB_CHOL_SHARE=8"; do
echo "== $cfg"
env $cfg timeout 300 python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
s=d['stage_ms_per_update']
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {k:s[k] for k in ('tracks','side_tallchol_slam_cols','side_slam_part','chol_gram','tallchol')})"
done
