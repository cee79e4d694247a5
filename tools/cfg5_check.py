"""BASELINE cfg-5 stress shape (50-pose window, 800 SLAM + 3200 MSCKF tracks, N = 2715): runs a few updates on the
device, checks covariance symmetry / positive semi-definiteness and state sanity, prints the per-stage times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from x_multi_agent_b200 import Filter
from x_multi_agent_b200.synth import Scenario, SynthConfig, record, replay

M, F, K = 50, 800, 3200
cfg = SynthConfig(M=M, F=F, K=60, seed=0, slam_init_frame=M, slam_lm_seed=7)
scn = Scenario(cfg)
flt = Filter(M, F, max_tracks=K, n_slots=64)
t0 = time.time()
replay(record(scn, M + 3), flt)
print("fill done", time.time() - t0, "s; n_poses", flt.n_poses, "n_features", flt.n_features)
scn.c.K = K
c = scn.c
fed = (M + 2) * c.imu_per_frame + c.latency_imu
flt.profile(True)
for k in range(M + 3, M + 8):
    upto = k * c.imu_per_frame + c.latency_imu
    for i in range(fed + 1, upto + 1):
        t = i * scn.dt_imu
        flt.process_imu(t, i, *scn.imu_sample(t), want_state=False)
    fed = upto
    m = scn.measurement(k)
    flt.set_measurement(m)
    t1 = time.perf_counter()
    st = flt.process_update_measurement()
    dt = time.perf_counter() - t1
    n = len(m.msckf_trks)
    inl = flt.debug_int("inlier0", n)
    p_true = scn.pose(st.time)[0]
    print(f"frame {k}: {dt*1e3:.2f} ms wall, msckf inliers {inl.sum()}/{n}, |p - p_true| = {np.linalg.norm(st.p - p_true):.3f} m")
prof = flt.profile_read()
print({k: round(v[0] / max(v[1], 1), 3) for k, v in prof.items() if v[1]})
P = flt.get_covariance()
ev = np.linalg.eigvalsh(0.5 * (P + P.T))
print("N", P.shape[0], "P asym", np.abs(P - P.T).max(), "lambda_min/max", ev.min() / ev.max(), "finite", np.isfinite(P).all())
