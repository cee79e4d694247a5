timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^ok" | tail -30
timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['gate_inlier_frac_last_step'], d['gpu_launches'])
print(d['stage_ms_per_update'])"
timeout 400 python tools/cfg5_check.py 2>&1 | tail -4
