"""Device time per Ekf::processImu call (fused single launch vs the two-kernel form), 2000 samples back to back."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import torch
from x_multi_agent_b200 import Filter
from x_multi_agent_b200.synth import Scenario, SynthConfig

M, F = 30, 200
cfg = SynthConfig(M=M, F=F, K=0, seed=3)
scn = Scenario(cfg)
s0 = scn.initial_state()
for mode in ("fused", "two"):
    if mode == "two":
        os.environ["XB_NO_FUSED_IMU"] = "1"
    dev = Filter(M, F, n_slots=250)
    st = torch.cuda.Stream()
    dev.set_stream(st.cuda_stream)
    dev.initialize_from_state(s0)
    samples = [(0.0, 0, *scn.imu_sample(0.0))] + [(i * 0.005, i, *scn.imu_sample(i * 0.005)) for i in range(1, 2101)]
    for s in samples[:100]:
        dev.process_imu(*s, want_state=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev.synchronize()
    e0.record(st)
    for s in samples[100:2100]:
        dev.process_imu(*s, want_state=False)
    e1.record(st)
    dev.synchronize()
    print(mode, "us per sample: %.2f" % (e0.elapsed_time(e1) * 1e3 / 2000), "phase clocks", dev.debug("FQ", 470)[460:465])
    dev.close()
