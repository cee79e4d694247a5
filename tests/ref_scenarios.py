"""Seeded scenarios shared by the golden-vector generator (oracle/tools/make_ref_golden.py), the CPU pinning tests
(oracle vs the reference compiled in place) and the GPU parity tests (CUDA path vs the reference).  Test
infrastructure only."""
import numpy as np

from x_multi_agent_b200.synth import Scenario, SynthConfig, record

# name -> (SynthConfig kwargs, frames, iekf_iter)
SCENARIOS = {
    "slam_msckf_short_churn": (dict(M=6, F=6, K=12, seed=1, n_short=2, churn=1), 14, 1),
    "cfg1_msckf_only": (dict(M=10, F=0, K=50, seed=0), 14, 1),
    "slam_only": (dict(M=5, F=4, K=0, seed=3), 12, 1),
    "early_slam_init_iekf2": (dict(M=8, F=8, K=10, seed=5, n_short=3, churn=2, slam_init_frame=3), 16, 2),
    # six consecutive updates without any measurement row while the window fills (K = 0): the reference's covariance
    # stays unsymmetrised over several clones (updater.cpp:106, state_manager.cpp:273-349, propagator.cpp:197-203)
    "consecutive_empty_updates": (dict(M=6, F=6, K=0, seed=1, churn=1), 14, 1),
}


def events(name):
    kw, frames, iekf = SCENARIOS[name]
    cfg = SynthConfig(**kw)
    return cfg, record(Scenario(cfg), frames), iekf


def state_rows(states):
    """Stack the estimates of a list of states (xvec layout, columns 0:31 + arrays) into one array."""
    return np.vstack([np.asarray(s.x if hasattr(s, "x") else s, dtype=np.float64) for s in states])


def oracle_xvec(s, M, F):
    """oracle.State -> xvec (include/xb200.h layout)."""
    x = np.zeros(32 + 7 * M + 3 * F)
    x[0:3], x[3:6], x[6:10], x[10:13], x[13:16] = s.p, s.v, s.q, s.b_w, s.b_a
    x[16:20], x[20:23], x[23:26], x[26:29] = s.q_ic, s.p_ic, s.w_m, s.a_m
    x[29], x[30] = s.time, s.seq
    x[32:32 + 3 * M] = s.p_array
    x[32 + 3 * M:32 + 7 * M] = s.q_array
    x[32 + 7 * M:32 + 7 * M + 3 * F] = s.f_array
    return x
