"""Timeline of the dataflow tile Cholesky on one steady-state cfg-2 update (set XB_CHOL_TRACE=1)."""
import os, sys
os.environ["XB_CHOL_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from x_multi_agent_b200 import Filter
from x_multi_agent_b200.synth import replay

scn, fill = bench.build_scenario(0)
flt = Filter(30, 200, max_tracks=800, n_slots=250, oc_projection=0, **bench.IMU_NOISE)
replay(fill, flt)
ev = bench.steady_events(scn, bench.N_FILL, 12)
for imu, m in ev:
    for (t, i, w, a) in imu:
        flt.process_imu(t, i, w, a, want_state=False)
    flt.set_measurement(m)
    flt.process_update_measurement()
tr = flt.debug("chol_trace", 10 * 6).reshape(-1, 10)   # the slab-column launch of the update: 6 tile columns
print("critical-path CTA, per tile column (us since start of the launch):")
print("col | start  tiles_loaded  syrk_done | warp 0: potrf_done  solve_done | warps 1-3: E_updated  D_published | column_end")
for r in tr:
    f = lambda k: "%8.2f" % (r[k] / 1e3)
    print(f"{int(r[0]):3d} |", f(2), f(9), f(3), "|", f(6), f(8), "|", f(7), f(4), "|", f(5))
