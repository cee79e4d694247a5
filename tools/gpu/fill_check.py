"""Warm-up sequence of bench.py (cfg-2 fill + a few steady-state updates) as a stand-alone script, for
compute-sanitizer runs: compute-sanitizer --tool memcheck python tools/gpu/fill_check.py"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import bench
from x_multi_agent_b200 import Filter
from x_multi_agent_b200.synth import replay

oc = int(sys.argv[1]) if len(sys.argv) > 1 else 0
scn, fill = bench.build_scenario(seed=0)
flt = Filter(30, 200, max_tracks=800, n_slots=250, oc_projection=oc)


def cb(k, m, st):
    flt.synchronize()
    print(k, len(m.msckf_trks), len(m.slam_trks), len(m.new_msckf_slam_trks), len(m.new_slam_std_trks), flush=True)


replay(fill, flt, cb)
ev = bench.steady_events(scn, bench.N_FILL, 3)
for imu, m in ev:
    for (t, seq, w, a) in imu:
        flt.process_imu(t, seq, w, a, want_state=False)
    flt.set_measurement(m)
    st = flt.process_update_measurement()
    print("steady", st.time, int(flt.debug_int("inlier0", 800).sum()), flush=True)
flt.close()
