#!/usr/bin/env python
"""Generates tests/golden/ref_sequences.npz by running the reference ITSELF (compiled in place by
oracle/ref_build/build_ref.sh into oracle/_ref/libxref.so) on the seeded scenarios of tests/ref_scenarios.py:
the state returned by every Ekf::processUpdateMeasurement and the newest ring-buffer state + covariance at the end.
Run from the repository root, in a container where /root/reference exists:

    bash oracle/ref_build/build_ref.sh && python oracle/tools/make_ref_golden.py

Test infrastructure only.  The fixture lets the GPU box (no /root/reference) and a tree without oracle/_ref check
the numpy oracle and the CUDA path against the reference's own outputs."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, os.fspath(ROOT))
sys.path.insert(0, os.fspath(ROOT / "tests"))

from oracle import refcpp  # noqa: E402
from ref_scenarios import SCENARIOS, events, state_rows  # noqa: E402
from x_multi_agent_b200.synth import replay  # noqa: E402


def main():
    out = {}
    for name in SCENARIOS:
        cfg, ev, iekf = events(name)
        ref = refcpp.RefFilter(cfg.M, cfg.F, sigma_img=cfg.sigma_img, n_slots=64, iekf_iter=iekf)
        states = []
        replay(ev, ref, lambda k, m, st: states.append(st.x.copy()))
        newest = ref.newest()
        n_p, n_f, anchors = ref.sm_info()
        out[f"{name}/updates"] = state_rows(states)
        out[f"{name}/newest_x"] = newest.x
        out[f"{name}/newest_cov"] = newest.cov
        out[f"{name}/sm"] = np.array([n_p, n_f] + anchors, dtype=np.int64)
        ref.close()
        print(f"{name}: {len(states)} updates, N = {newest.cov.shape[0]}, "
              f"max |P - P^T| = {np.abs(newest.cov - newest.cov.T).max():.2e}")
    dst = ROOT / "tests" / "golden" / "ref_sequences.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
