"""The C++ host binding (include/x/xb200_binding.hpp: x::Ekf / x::VioUpdater / x::State over the C ABI) replays the
same event stream as the ctypes path and must produce identical states."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from x_multi_agent_b200 import Filter
from x_multi_agent_b200.synth import Scenario, SynthConfig, record, replay

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _serialise(cfg, ev, max_tracks):
    out = [float(cfg.M), float(cfg.F), float(max_tracks), cfg.sigma_img]
    for e in ev:
        if e[0] == "init":
            out.append(0.0)
            out += list(e[1].x) + list(np.asarray(e[1].cov).ravel())
        elif e[0] == "imu":
            out += [1.0, e[1], float(e[2]), *e[3], *e[4]]
        else:
            m = e[1]
            out += [2.0, m.timestamp]
            for tl in (m.slam_trks, m.msckf_trks, m.msckf_short_trks, m.new_slam_std_trks, m.new_msckf_slam_trks):
                out.append(float(len(tl)))
                for t in tl:
                    t = np.asarray(t)
                    out.append(float(t.shape[0]))
                    out += list(t.ravel())
            out.append(float(len(m.lost_slam_trk_idxs)))
            out += [float(i) for i in m.lost_slam_trk_idxs]
            if m.range is not None:
                out += [1.0, m.range.timestamp, m.range.range, *m.range.img_pt_n, *[float(i) for i in m.range.tr_feat_ids]]
            else:
                out.append(0.0)
            if m.sun_angle is not None:
                out += [1.0, m.sun_angle.timestamp, m.sun_angle.x_angle, m.sun_angle.y_angle]
            else:
                out.append(0.0)
    return np.asarray(out, dtype=np.float64)


def test_cxx_binding_replays_identically(tmp_path):
    cfg = SynthConfig(M=6, F=6, K=12, seed=2, n_short=2, churn=1, range_every=2, sun_every=3)  # + LRF and sun-sensor rows
    ev = record(Scenario(cfg), 12)
    exe = tmp_path / "test_x_api"
    libdir = ROOT / "x_multi_agent_b200"
    # Eigen is not installed here: the binding is compiled against the stand-in of the test infrastructure
    subprocess.run(["g++", "-std=c++17", "-O2", f"-I{ROOT / 'include'}", f"-I{ROOT / 'oracle' / 'ref_build' / 'shim'}", "-o",
                    os.fspath(exe),
                    os.fspath(ROOT / "tests" / "cxx" / "test_x_api.cpp"), f"-L{libdir}", "-lxb200", f"-Wl,-rpath,{libdir}"],
                   check=True)
    _serialise(cfg, ev, 16).tofile(tmp_path / "events.bin")
    r = subprocess.run([os.fspath(exe), os.fspath(tmp_path / "events.bin"), os.fspath(tmp_path / "out.bin")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "out.bin")
    dev = Filter(cfg.M, cfg.F, max_tracks=16, n_slots=64, sigma_img=cfg.sigma_img, sigma_range=0.05)
    states = []
    replay(ev, dev, lambda k, m, st: states.append(st.x.copy()))
    states.append(dev.get_state().x)
    want = np.concatenate(states)
    assert got.shape == want.shape
    assert np.array_equal(got, want), np.abs(got - want).max()
    dev.close()
