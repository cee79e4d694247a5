for cfg in "XB_SIDE_PRIORITY=0" "XB_X=1"; do
echo "== $cfg"
env $cfg timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print(d['stage_ms_per_update'])"
done
