"""F_d / Q_d of one IMU step: fused single-launch kernel against the two-kernel form (XB_NO_FUSED_IMU=1)."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
from x_multi_agent_b200 import Filter
from x_multi_agent_b200.synth import Scenario, SynthConfig

cfg = SynthConfig(M=4, F=3, K=0, seed=3)
scn = Scenario(cfg)
s0 = scn.initial_state()
out = {}
for mode in ("fused", "two"):
    if mode == "two":
        os.environ["XB_NO_FUSED_IMU"] = "1"
    dev = Filter(cfg.M, cfg.F, n_slots=64)
    dev.initialize_from_state(s0)
    for (t, seq, w, a) in [(0.0, 0, *Scenario(cfg).imu_sample(0.0))] + Scenario(cfg).imu_between(0, 1)[:2]:
        dev.process_imu(t, seq, w, a)
    out[mode] = dev.debug("FQ", 450).copy()
    dev.close()
dF = np.abs(out["fused"][:225] - out["two"][:225]).reshape(15, 15)
dQ = np.abs(out["fused"][225:] - out["two"][225:]).reshape(15, 15)
print("max dF", dF.max(), "max dQ", dQ.max())
print("dQ nonzero at", np.argwhere(dQ > 1e-18)[:20].tolist())
print("Q fused", out["fused"][225:].reshape(15, 15)[np.nonzero(dQ > 1e-18)][:10], "two", out["two"][225:].reshape(15, 15)[np.nonzero(dQ > 1e-18)][:10])
print("max dF", dF.max())
from oracle.qd_poly import qd_poly
from oracle.quat import rot_raw
dev = Filter(cfg.M, cfg.F, n_slots=64)
dev.initialize_from_state(s0)
samples = [(0.0, 0, *Scenario(cfg).imu_sample(0.0))] + Scenario(cfg).imu_between(0, 1)[:2]
sts = [dev.process_imu(*s) for s in samples]
x0, x1 = sts[1].x, sts[2].x
dt = x1[29] - x0[29]
w1, a1 = x1[23:26] - x1[10:13], x1[26:29] - x1[13:16]
Qo = qd_poly(dt, rot_raw(x1[6:10]), w1, a1, 0.0083, 0.00083, 0.0013, 0.00013)
Qf = dev.debug("FQ", 450)[225:].reshape(15, 15)
print("fused vs oracle Q", np.abs(Qf - Qo).max(), "two vs oracle", np.abs(out["two"][225:].reshape(15, 15) - Qo).max())
Qw0 = qd_poly(dt, rot_raw(x1[6:10]), w1 * 0, a1 * 0, 0.0083, 0.00083, 0.0013, 0.00013)
print("fused vs oracle Q with w=a=0", np.abs(Qf - Qw0).max())
Qc = qd_poly(dt, np.eye(3), w1, a1, 0.0083, 0.00083, 0.0013, 0.00013)
print("fused vs oracle Q with C=I", np.abs(Qf - Qc).max())
Qfu = out["fused"][225:].reshape(15, 15)
print("FUSED vs oracle Q", np.abs(Qfu - Qo).max(), "w=a=0:", np.abs(Qfu - Qw0).max(), "C=I:", np.abs(Qfu - Qc).max())
for name, (ww, aa, CC) in {"a=0": (w1, a1 * 0, rot_raw(x1[6:10])), "w=0": (w1 * 0, a1, rot_raw(x1[6:10])),
                           "C(q0)": (w1, a1, rot_raw(x0[6:10])), "prev sample": (x0[23:26] - x0[10:13], x0[26:29] - x0[13:16], rot_raw(x0[6:10]))}.items():
    print(name, np.abs(Qfu - qd_poly(dt, CC, ww, aa, 0.0083, 0.00083, 0.0013, 0.00013)).max())
bad = np.argwhere(np.abs(Qfu - Qo) > 1e-15)
print("entries off:", len(bad), "of", int((Qo != 0).sum()), bad[:12].tolist())
