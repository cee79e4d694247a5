"""Two cfg-2 agents in one process (seeds 0 / 1, shared SLAM landmark set): SLAM-SLAM CI gates of agent 0 against agent 1's payload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from x_multi_agent_b200 import Filter
from x_multi_agent_b200.synth import replay

flts, scns = [], []
for seed in (0, 1):
    scn, fill = bench.build_scenario(seed)
    f = Filter(30, 200, max_tracks=800, n_slots=250, sigma_landmark=1.0, ci_slam_w=0.1, ci_msckf_w=0.1)
    replay(fill, f)
    flts.append(f); scns.append(scn)
PL = flts[0].ci_payload_len()
gathered = torch.zeros((2, PL), dtype=torch.float64, device="cuda")
for a in range(2):
    flts[a].ci_pack(gathered[a].data_ptr())
    flts[a].synchronize()
g = gathered.cpu().numpy()
s0, s1 = flts[0].get_state(), flts[1].get_state()
print("agent positions", s0.p, s1.p, "times", s0.time, s1.time)
for a in range(2):
    pay = g[a, 8:].reshape(-1, 13)
    print("agent", a, "valid", int(pay[:, 0].sum()), "world pos of features 0..2:", pay[:3, 1:4].round(3).tolist())
print("true landmarks 0..2 (agent 0 scenario):", scns[0].slam_lm[:3].round(3).tolist())
print("true landmarks 0..2 (agent 1 scenario):", scns[1].slam_lm[:3].round(3).tolist())
print("feat_lm agent0[:6]", scns[0].feat_lm[:6], "agent1[:6]", scns[1].feat_lm[:6])
matches = [(1, f, f) for f in range(200)]
flts[0].process_others_packed(s0.time, gathered.data_ptr(), 2, matches, want_state=False)
flts[0].synchronize()
gates = flts[0].ci_last_gates(200)
print("inlier frac", gates[:, 0].mean(), "gamma quantiles", np.nanquantile(gates[:, 1], [0.1, 0.5, 0.9]))
