"""End to end from feature MATCHES (SURVEY 8 rows f-1 / f-2 in front of the hot path): x_multi_agent_b200.VIO
(YAML-style parameters -> camera + track manager on the host -> device filter) against the same pipeline assembled from
the oracle pieces (oracle.track_manager + the fp64 oracle filter), on one synthetic stream of 10-double match vectors
(VIO::importMatches format, vio.cpp:372-434) and IMU samples."""
import numpy as np
import pytest

from x_multi_agent_b200.synth import Scenario, SynthConfig

pytestmark = pytest.mark.gpu

NL = 700   # landmarks on the ground patch under the trajectory
PARAMS = dict(cam1_fx=0.46, cam1_fy=0.61, cam1_cx=0.5, cam1_cy=0.5, cam1_s=0.0, cam1_img_width=640, cam1_img_height=480,
              sigma_img=0.6 / 294.4, n_tiles_h=2, n_tiles_w=2, msckf_baseline=12.0, min_track_length=5, rho_0=0.3, sigma_rho_0=0.3,
              iekf_iter=1, n_poses_max=8, n_slam_features_max=6, state_buffer_size=120,
              sigma_dp=[0.1, 0.1, 0.1], sigma_dv=[0.1, 0.1, 0.1], sigma_dtheta=[2, 2, 2], sigma_dbw=[0.5, 0.5, 0.5],
              sigma_dba=[0.05, 0.05, 0.05], n_w=0.0083, n_bw=0.00083, n_a=0.0013, n_ba=0.00013)   # the scenario's IMU


def _stream(seed, frames):
    """IMU samples and match vectors of a camera looking down at a patch of landmarks."""
    scn = Scenario(SynthConfig(M=8, F=6, K=0, seed=seed, height=4.0))
    rng = np.random.default_rng(seed + 7)
    lms = np.column_stack([rng.uniform(-8, 8, NL), rng.uniform(-6, 6, NL), rng.uniform(-0.6, 0.6, NL)])
    fx, fy, cx, cy = 0.46 * 640, 0.61 * 480, 320.0, 240.0
    last_px = {}
    events = []
    fed = 0
    for k in range(frames):
        upto = k * scn.c.imu_per_frame + scn.c.latency_imu
        for i in range(fed + 1, upto + 1):
            t = i * scn.dt_imu
            events.append(("imu", t, i, *scn.imu_sample(t)))
        fed = max(fed, upto)
        tk = scn.frame_time(k)
        pc, Rc = scn.cam_pose(tk)
        x = (lms - pc) @ Rc
        px = np.column_stack([x[:, 0] / x[:, 2] * fx + cx, x[:, 1] / x[:, 2] * fy + cy]) + rng.normal(0, 0.3, (len(lms), 2))
        vis = (x[:, 2] > 1.0) & (px[:, 0] > 8) & (px[:, 0] < 632) & (px[:, 1] > 8) & (px[:, 1] < 472) & (rng.random(len(lms)) > 0.03)
        rows = []
        cur = {}
        for j in np.flatnonzero(vis):
            cur[j] = px[j]
            if j in last_px:
                rows.append([0, scn.frame_time(k - 1), *last_px[j], tk, *px[j], 0, 0, 0])
        last_px = cur
        order = rng.permutation(len(rows))
        events.append(("matches", tk, k, np.array(rows, dtype=float).reshape(-1, 10)[order]))
    return scn, events


class _OracleVIO:
    """The same facade assembled from the oracle pieces (test infrastructure)."""

    def __init__(self, vio):
        from oracle.track_manager import TrackManagerOracle
        from oracle_driver import OracleFilter
        import oracle
        p = vio.params
        self.p = p
        bx = p["msckf_baseline"] / (p["cam1_img_width"] * p["cam1_fx"])
        by = p["msckf_baseline"] / (p["cam1_img_height"] * p["cam1_fy"])
        self.tm = TrackManagerOracle(p["cam1_fx"], p["cam1_fy"], p["cam1_cx"], p["cam1_cy"], p["cam1_s"], p["cam1_img_width"],
                                     p["cam1_img_height"], bx, by, p["n_tiles_h"], p["n_tiles_w"])
        noise = oracle.ImuNoise()
        noise.n_w, noise.n_bw, noise.n_a, noise.n_ba = p["n_w"], p["n_bw"], p["n_a"], p["n_ba"]
        self.f = OracleFilter(p["n_poses_max"], p["n_slam_features_max"], sigma_img=p["sigma_img"], rho_0=p["rho_0"],
                              sigma_rho_0=p["sigma_rho_0"], iekf_iter=p["iekf_iter"], n_slots=p["state_buffer_size"],
                              g=tuple(p["g"]), noise=noise)
        self.vio = vio

    def process_matches(self, t, mv):
        from x_multi_agent_b200.filter import Measurement
        p, ekf = self.p, self.f.ekf
        n_poses = self.f.upd.sm.n_poses
        if n_poses == 0:
            mv = mv[:0]
        idx = ekf.buf.closest_idx(t)
        s = ekf.buf.states[idx]
        M = p["n_poses_max"]
        qa = np.asarray(s.q_array).reshape(M, 4)[:n_poses]
        size_out = min(M - 1, n_poses)
        from x_multi_agent_b200.vio import _qmul
        rots = np.vstack([qa[n_poses - size_out:], _qmul(s.q / np.linalg.norm(s.q), s.q_ic / np.linalg.norm(s.q_ic))[None]])
        self.tm.manage_tracks(mv, rots, M, p["n_slam_features_max"], p["min_track_length"])

        def tl(which, so=0):
            off, xy = self.tm.get_list(which, so)
            return [xy[off[i]:off[i + 1]].copy() for i in range(len(off) - 1)]
        m = Measurement(t, tl(4, M), tl(0), tl(1), tl(2), tl(3), [int(i) for i in self.tm.lost])
        self._sensors(m)
        self.f.set_measurement(m)
        return self.f.process_update_measurement()

    last_range = None   # (timestamp, range) as VIO::setLastRangeMeasurement keeps it
    last_sun = None

    def _sensors(self, m):
        """vio.cpp:288-298 + vio_updater.cpp:358-369 from independent pieces: the Delaunay facet around the hard-coded image
        point (320.5, 240.5) through qhull, the normalised LRF image point through the oracle camera."""
        from oracle.track_manager import Feat
        from scipy.spatial import Delaunay
        from x_multi_agent_b200.filter import RangeMeasurement
        if self.last_range is not None and self.last_range[0] > 0.1 and m.slam_trks and len(self.tm.slam) >= 3:
            pts = np.array([[t[-1].xd, t[-1].yd] for t in self.tm.slam], dtype=np.float32).astype(np.float64)
            tri = Delaunay(pts)
            sx = tri.find_simplex(np.array([[320.5, 240.5]]))[0]
            if sx >= 0:
                f = Feat(xd=(self.p["cam1_img_width"] + 1) / 2.0, yd=(self.p["cam1_img_height"] + 1) / 2.0)
                self.tm.undistort(f)
                pt = (f.x * self.tm.inv_fx - self.tm.cx_n, f.y * self.tm.inv_fy - self.tm.cy_n)
                m.range = RangeMeasurement(self.last_range[0], self.last_range[1], pt, sorted(int(i) for i in tri.simplices[sx]))
                self.n_range = getattr(self, "n_range", 0) + 1
        if self.last_sun is not None:
            m.sun_angle = self.last_sun


def test_vio_facade_from_match_vectors_matches_the_oracle_pipeline():
    from x_multi_agent_b200 import VIO
    from test_gpu_parity import Report, compare_state
    scn, events = _stream(11, 40)
    vio = VIO()
    params = dict(PARAMS)
    s0 = scn.initial_state()
    params.update(p=list(s0.p), v=list(s0.v), q=[s0.q[3], s0.q[0], s0.q[1], s0.q[2]], b_w=list(s0.b_w), b_a=list(s0.b_a),
                  cam1_p_ic=list(scn.p_ic), cam1_q_ic=[scn.q_ic[3], scn.q_ic[0], scn.q_ic[1], scn.q_ic[2]])
    vio.set_up(params, max_tracks=256)
    vio.init_at_time(0.0)
    ora = _OracleVIO(vio)
    ora.f.initialize_from_state(vio.initial_state(0.0))
    n_upd, n_msckf, n_feat = 0, 0, 0
    rp = Report()
    last_d = last_o = None
    for ev in events:
        if ev[0] == "imu":
            vio.process_imu(*ev[1:])
            ora.f.process_imu(*ev[1:])
        else:
            _, t, k, mv = ev
            d = vio.process_matches_measurement(t, k, mv)
            o = ora.process_matches(t, mv)
            assert (d is None) == (o is None)
            if d is not None:
                n_upd += 1
                last_d, last_o = d, o
                n_msckf += len(vio.track_manager.get_list(0)[0]) - 1
                n_feat = vio.filter.n_features
    assert n_upd >= 35 and n_msckf > 20 and n_feat == params["n_slam_features_max"], (n_upd, n_msckf, n_feat)
    compare_state(rp, "last update", last_d, last_o, params["n_poses_max"], params["n_slam_features_max"], tol_scale=100.0,
                  cov=False)
    rp.done()
    # the estimate follows the truth (sanity of the whole chain, not a parity statement)
    p_true = scn.pose(events[-1][1])[0]
    assert np.linalg.norm(last_d.p - p_true) < 0.5
    vio.close()


def test_vio_facade_with_range_and_sun_measurements():
    """VIO::setLastRangeMeasurement / setLastSunAngleMeasurement (vio.cpp:217-224) through the facade: the LRF facet comes
    from the product's Delaunay lookup on one side and from qhull on the other, the rows from the device and from the
    oracle; the facet's vertex order is irrelevant to the row (range_update.cpp:141-242 is symmetric
    under permutations of the three features)."""
    from x_multi_agent_b200 import VIO
    from x_multi_agent_b200.filter import SunAngleMeasurement
    from test_gpu_parity import Report, compare_state
    scn, events = _stream(11, 30)
    vio = VIO()
    params = dict(PARAMS)
    s0 = scn.initial_state()
    params.update(p=list(s0.p), v=list(s0.v), q=[s0.q[3], s0.q[0], s0.q[1], s0.q[2]], b_w=list(s0.b_w), b_a=list(s0.b_a),
                  cam1_p_ic=list(scn.p_ic), cam1_q_ic=[scn.q_ic[3], scn.q_ic[0], scn.q_ic[1], scn.q_ic[2]], sigma_range=0.05)
    vio.set_up(params, max_tracks=256)
    vio.init_at_time(0.0)
    ora = _OracleVIO(vio)
    ora.f.upd.sigma_range = 0.05
    ora.f.initialize_from_state(vio.initial_state(0.0))
    rp = Report()
    last_d = last_o = None
    n_upd = 0
    for ev in events:
        if ev[0] == "imu":
            vio.process_imu(*ev[1:])
            ora.f.process_imu(*ev[1:])
        else:
            _, t, k, mv = ev
            # the camera looks down from ~4 m: a plausible altimeter reading, and a sun-sensor reading on every third frame
            pc, Rc = scn.cam_pose(t)
            rng_m = float(pc[2] / max(Rc[2, 2] * -1.0, 0.2)) if Rc[2, 2] < 0 else 4.0
            vio.set_last_range_measurement(t, rng_m)
            ora.last_range = (t, rng_m)
            if k % 3 == 0:
                sun = scn._sun_measurement(k)
                vio.set_last_sun_angle_measurement(sun.timestamp, sun.x_angle, sun.y_angle)
                ora.last_sun = SunAngleMeasurement(sun.timestamp, sun.x_angle, sun.y_angle)
            d = vio.process_matches_measurement(t, k, mv)
            o = ora.process_matches(t, mv)
            assert (d is None) == (o is None)
            if d is not None:
                n_upd += 1
                last_d, last_o = d, o
    assert n_upd >= 25 and getattr(ora, "n_range", 0) >= 5, (n_upd, getattr(ora, "n_range", 0))
    compare_state(rp, "last update", last_d, last_o, params["n_poses_max"], params["n_slam_features_max"], tol_scale=1000.0,
                  cov=False)
    rp.done()
    vio.close()


def _yaml(params):
    """The parameter file a caller of the reference writes (OpenCV FileStorage flavour: directive line, flow sequences)."""
    lines = ["%YAML:1.0", "---"]
    for k, v in params.items():
        lines.append(f"{k}: [{', '.join(repr(float(e)) for e in v)}]" if isinstance(v, (list, tuple)) else f"{k}: {v!r}")
    return "\n".join(lines) + "\n"


def test_cxx_vio_facade_equals_the_python_facade(tmp_path):
    """The C++ x::VIO (include/x/vio/vio.h: loadParamsFromYaml -> setUp -> initAtTime -> processImu /
    setLastRangeMeasurement / setLastSunAngleMeasurement / processMatchesMeasurement, the reference's public entry points
    vio.h:43-137) on the stream of the test above: every updated state equals the Python facade's -- both drive the same
    library, the C++ side through the reference's own template method Updater::update with the matches sorted into tracks
    inside VioUpdater::preProcess (vio_updater.cpp:142-179)."""
    import os
    import subprocess
    from pathlib import Path
    from x_multi_agent_b200 import VIO
    root = Path(__file__).resolve().parents[1]
    scn, events = _stream(11, 30)
    params = dict(PARAMS)
    s0 = scn.initial_state()
    params.update(p=list(s0.p), v=list(s0.v), q=[s0.q[3], s0.q[0], s0.q[1], s0.q[2]], b_w=list(s0.b_w), b_a=list(s0.b_a),
                  cam1_p_ic=list(scn.p_ic), cam1_q_ic=[scn.q_ic[3], scn.q_ic[0], scn.q_ic[1], scn.q_ic[2]], sigma_range=0.05,
                  g=[0.0, 0.0, -9.81], sigma_landmark=0.1, ci_slam_w=0.5, ci_msckf_w=0.5)   # the last three: MULTI_UAV build only
    (tmp_path / "params.yaml").write_text(_yaml(params))
    vio = VIO()
    vio.set_up(params, max_tracks=256)
    vio.init_at_time(0.0)
    out, want = [], []
    n_upd = 0
    for ev in events:
        if ev[0] == "imu":
            vio.process_imu(*ev[1:])
            out += [1.0, ev[1], float(ev[2]), *ev[3], *ev[4]]
        else:
            _, t, k, mv = ev
            pc, Rc = scn.cam_pose(t)
            rng_m = float(pc[2] / max(Rc[2, 2] * -1.0, 0.2)) if Rc[2, 2] < 0 else 4.0
            vio.set_last_range_measurement(t, rng_m)
            out += [3.0, t, rng_m]
            if k % 3 == 0:
                sun = scn._sun_measurement(k)
                vio.set_last_sun_angle_measurement(sun.timestamp, sun.x_angle, sun.y_angle)
                out += [4.0, sun.timestamp, sun.x_angle, sun.y_angle]
            d = vio.process_matches_measurement(t, k, mv)
            out += [2.0, t, float(k), float(len(mv)), *np.asarray(mv, dtype=float).ravel()]
            want.append(1.0 if d is not None else 0.0)
            want += list(d.x) if d is not None else [0.0] * vio.filter.LX
            n_upd += d is not None
    assert n_upd >= 25
    # SLAM features of the newest state in world coordinates (StateManager::computeSLAMCartesianFeaturesForState)
    flt = vio.filter
    s = flt.get_state()
    M, F = params["n_poses_max"], params["n_slam_features_max"]
    nf = flt.n_features
    anchors = flt.anchor_idxs[:nf]
    np.asarray(out, dtype=np.float64).tofile(tmp_path / "events.bin")
    exe = tmp_path / "test_x_vio"
    libdir = root / "x_multi_agent_b200"
    subprocess.run(["g++", "-std=c++17", "-O2", f"-I{root / 'include'}", f"-I{root / 'oracle' / 'ref_build' / 'shim'}", "-o",
                    os.fspath(exe), os.fspath(root / "tests" / "cxx" / "test_x_vio.cpp"), f"-L{libdir}", "-lxb200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([os.fspath(exe), os.fspath(tmp_path / "params.yaml"), os.fspath(tmp_path / "events.bin"),
                        os.fspath(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "out.bin")
    want = np.asarray(want)
    LX = flt.LX
    assert got.shape[0] >= want.shape[0] + 1
    g = got[:want.shape[0]].reshape(-1, LX + 1).copy()
    w = want.reshape(-1, LX + 1).copy()
    g[:, 1 + 31] = w[:, 1 + 31] = 0.0          # xvec[31] is reserved (not part of x::State)
    assert np.array_equal(g[:, 0], w[:, 0])
    err = np.abs(g - w).max()
    assert err <= 1e-10, err
    # world coordinates of the SLAM features, against the same formula on the Python side
    n_xyz = int(got[want.shape[0]])
    assert n_xyz == nf and nf == F
    xyz = got[want.shape[0] + 1:].reshape(n_xyz, 3)
    from oracle.quat import rot as q_rot
    for i in range(nf):
        a = int(anchors[i])
        qa = s.q_array.reshape(M, 4)[a]
        al, be, rho = s.f_array.reshape(F, 3)[i]
        ref = s.p_array.reshape(M, 3)[a] + q_rot(qa) @ np.array([al, be, 1.0]) / rho
        assert np.abs(xyz[i] - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    assert np.all(np.isfinite(xyz))
    vio.close()
    # -DMULTI_UAV flavour of the same program: two agents (two filters on this GPU) on the stream, VIO::getDataToSend on one,
    # VIO::processOtherMeasurements on the other with every SLAM feature matched to its twin: zero landmark residuals, so the
    # SLAM-SLAM covariance intersection moves no estimate and returns a finite covariance that differs from the prior one
    exe2 = tmp_path / "test_x_vio_multi"
    subprocess.run(["g++", "-std=c++17", "-O2", "-DMULTI_UAV", f"-I{root / 'include'}", f"-I{root / 'oracle' / 'ref_build' / 'shim'}",
                    "-o", os.fspath(exe2), os.fspath(root / "tests" / "cxx" / "test_x_vio.cpp"), f"-L{libdir}", "-lxb200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([os.fspath(exe2), os.fspath(tmp_path / "params.yaml"), os.fspath(tmp_path / "events.bin"),
                        os.fspath(tmp_path / "out_multi.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "multi ok" in r.stdout, r.stdout + r.stderr
