"""Phase breakdown of the per-track kernel on one steady-state cfg-2 update (XB_TRACK_PROF=1): mean/max SM clocks
between the phase boundaries of k_tracks, over all tracks."""
import os, sys
os.environ["XB_TRACK_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from x_multi_agent_b200 import Filter
from x_multi_agent_b200.synth import replay

scn, fill = bench.build_scenario(0)
flt = Filter(30, 200, max_tracks=800, n_slots=250, oc_projection=0, **bench.IMU_NOISE)
replay(fill, flt)
ev = bench.steady_events(scn, bench.N_FILL, 3)
for imu, m in ev:
    for (t, i, w, a) in imu:
        flt.process_imu(t, i, w, a, want_state=False)
    flt.set_measurement(m)
    flt.process_update_measurement()
tr = flt.debug("track_prof", 12 * 800).reshape(-1, 12)[:, :11]
names = ["dlt", "gauss-newton", "jacobians", "mgs", "B=U^T J", "X=JPJ^T", "Y,Z,V,S,aug", "cholesky", "woodbury", "outputs"]
d = np.diff(tr, axis=1)
print(f"tracks {len(tr)}  total cycles mean {tr[:, 10].mean():.0f} max {tr[:, 10].max():.0f}")
for k, n in enumerate(names):
    print(f"  {n:14s} mean {d[:, k].mean():9.0f}  max {d[:, k].max():9.0f}  ({100 * d[:, k].mean() / tr[:, 10].mean():5.1f} %)")
