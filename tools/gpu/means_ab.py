"""A/B of the re-propagation means kernel (XB_OLD_MEANS toggled per filter): first frame where the newest state differs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from x_multi_agent_b200 import Filter
from x_multi_agent_b200.synth import Scenario, SynthConfig, record, replay

cfg = SynthConfig(M=30, F=200, K=96, seed=3, slam_init_frame=30, churn=3, n_short=4)
ev = record(Scenario(cfg), 40)
outs = []
for old in (1, 0):
    if old: os.environ["XB_OLD_MEANS"] = "1"
    else: os.environ.pop("XB_OLD_MEANS", None)
    f = Filter(cfg.M, cfg.F, max_tracks=cfg.K, sigma_img=cfg.sigma_img, n_slots=64)
    res = []
    def on_upd(k, m, st):
        res.append((np.array(st.x), np.array(f.get_state().x)))
    replay(ev, f, on_upd)
    outs.append(res)
    f.close()
for k in range(len(outs[0])):
    du = np.abs(outs[0][k][0] - outs[1][k][0]).max()
    dn = np.abs(outs[0][k][1] - outs[1][k][1])
    print(k, "upd diff %.2e newest diff %.2e at %d" % (du, dn.max(), int(dn.argmax())))
    if dn.max() > 1e-10:
        print(" first 32 newest diff:", np.array2string(dn[:32], precision=1))
        break
