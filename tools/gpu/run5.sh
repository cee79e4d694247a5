timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^ok" | tail -30
timeout 300 python tools/track_prof.py 2>&1 | tail -14
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['gate_inlier_frac_last_step'], d['gpu_launches'])
print(d['stage_ms_per_update'])"


ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s2_launches.csv k_assemble 6 2>&1 | tail -50
