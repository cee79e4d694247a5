"""Synthetic IMU + feature-track generator at the VioUpdater::preProcess seam.

numpy only.  Produces, for a seeded 6-DoF trajectory, the IMU stream and -- for every camera frame --
the five track lists + lost-feature indexes that the reference's front end would hand to the filter
(src/x/vio/vio_updater.cpp:172-179).  Observations are pinhole projections of fixed landmarks from the
TRUE camera poses plus Gaussian pixel noise and a fraction of gross outliers.
"""
from dataclasses import dataclass

import numpy as np

from .filter import Measurement, RangeMeasurement, State, SunAngleMeasurement, K_CORE


def _rx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _ry(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def rot_to_quat(R):
    """Rotation matrix -> (x,y,z,w)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q / np.linalg.norm(q)


def quat_to_rot(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


@dataclass
class SynthConfig:
    M: int = 10                 # n_poses_max
    F: int = 0                  # SLAM features
    K: int = 50                 # MSCKF tracks per update
    seed: int = 0
    cam_rate: float = 20.0
    imu_per_frame: int = 10
    latency_imu: int = 10       # IMU samples already buffered after the frame when its update arrives
    sigma_img: float = 1.0 / 320.0
    outlier_frac: float = 0.02
    radius: float = 3.0
    freq: float = 0.05
    height: float = 10.0
    n_short: int = 0            # short MSCKF tracks per update (length 2..window-1)
    slam_init_frame: int = -1   # frame at which SLAM features are initialised (-1: when the window is full)
    slam_msckf_init_frac: float = 0.5   # fraction initialised through MSCKF-SLAM promotion (rest: standard)
    churn: int = 0              # SLAM features lost (and re-added) per update once running
    slam_lm_seed: int = -1      # >= 0: SLAM landmark set drawn from this seed (shared by all agents of a CI scenario)
    n_w: float = 0.0083
    n_bw: float = 0.00083
    n_a: float = 0.0013
    n_ba: float = 0.00013
    init_err_scale: float = 0.5  # initial estimation error = N(0,1) * sigma * init_err_scale (sigma = initial P)
    range_every: int = 0        # > 0: a laser-range measurement on every range_every-th frame with >= 3 SLAM tracks
    sigma_range: float = 0.05
    sun_every: int = 0          # > 0: a sun-sensor measurement on every sun_every-th frame
    sun_noise_deg: float = 1.0


class Scenario:
    G = np.array([0.0, 0.0, -9.81])

    def __init__(self, cfg: SynthConfig):
        self.c = cfg
        self.rng = np.random.Generator(np.random.PCG64(cfg.seed))
        self.dt_imu = 1.0 / (cfg.cam_rate * cfg.imu_per_frame)
        self.R_ic = _rx(np.pi)
        self.q_ic = rot_to_quat(self.R_ic)
        self.p_ic = np.array([0.05, 0.0, -0.02])
        self.b_w = self.rng.normal(0, 0.01, 3)
        self.b_a = self.rng.normal(0, 0.01, 3)
        self.phase = self.rng.uniform(0, 2 * np.pi, 4)
        # SLAM landmarks: a patch under the trajectory centre that stays in view
        n_l = max(cfg.F * 2 + 8, 8)
        lrng = self.rng if cfg.slam_lm_seed < 0 else np.random.Generator(np.random.PCG64(cfg.slam_lm_seed))
        self.slam_lm = np.column_stack([lrng.uniform(-2.0, 2.0, n_l), lrng.uniform(-1.5, 1.5, n_l),
                                        lrng.uniform(-2.0, 0.5, n_l)])
        self.next_lm = 0
        self.feat_lm = []           # landmark id per active SLAM feature slot
        self.feat_obs = []          # observation history per active SLAM feature
        self.imu_seq = 0
        self._cam_cache = {}
        self._last_lms = np.zeros((0, 3))
        self.last_msckf_lms, self.last_short_lms = np.zeros((0, 3)), []

    # ---- truth ---------------------------------------------------------------------------------------
    def pose(self, t):
        """(p, R, v, a, w_body) of the IMU frame at time t (analytic)."""
        c = self.c
        w0 = 2 * np.pi * c.freq
        ph = self.phase
        r = c.radius
        p = np.array([r * np.sin(w0 * t + ph[0]), r * np.sin(2 * w0 * t + ph[1]) * 0.6, c.height + 0.5 * np.sin(1.5 * w0 * t + ph[2])])
        v = np.array([r * w0 * np.cos(w0 * t + ph[0]), 1.2 * r * w0 * np.cos(2 * w0 * t + ph[1]),
                      0.75 * w0 * np.cos(1.5 * w0 * t + ph[2])])
        a = np.array([-r * w0 ** 2 * np.sin(w0 * t + ph[0]), -2.4 * r * w0 ** 2 * np.sin(2 * w0 * t + ph[1]),
                      -1.125 * w0 ** 2 * np.sin(1.5 * w0 * t + ph[2])])
        yaw, dyaw = 0.4 * np.sin(w0 * t + ph[3]), 0.4 * w0 * np.cos(w0 * t + ph[3])
        pit, dpit = 0.08 * np.sin(2.3 * w0 * t), 0.08 * 2.3 * w0 * np.cos(2.3 * w0 * t)
        rol, drol = 0.08 * np.cos(1.7 * w0 * t), -0.08 * 1.7 * w0 * np.sin(1.7 * w0 * t)
        R = _rz(yaw) @ _ry(pit) @ _rx(rol)
        w = np.array([drol - dyaw * np.sin(pit), dpit * np.cos(rol) + dyaw * np.sin(rol) * np.cos(pit),
                      -dpit * np.sin(rol) + dyaw * np.cos(rol) * np.cos(pit)])
        return p, R, v, a, w

    def cam_pose(self, t):
        key = round(t * 1e9)
        hit = self._cam_cache.get(key)
        if hit is None:
            p, R, *_ = self.pose(t)
            hit = (p + R @ self.p_ic, R @ self.R_ic)
            self._cam_cache[key] = hit
        return hit

    def frame_time(self, k):
        return k / self.c.cam_rate

    def imu_sample(self, t):
        c = self.c
        p, R, v, a, w = self.pose(t)
        sd = 1.0 / np.sqrt(self.dt_imu)
        w_m = w + self.b_w + self.rng.normal(0, c.n_w * sd, 3)
        a_m = R.T @ (a - self.G) + self.b_a + self.rng.normal(0, c.n_a * sd, 3)
        return w_m, a_m

    # ---- initial state (VIO::initAtTime, vio.cpp:54-111) ------------------------------------------------
    def initial_state(self, sigmas=(0.1, 0.1, 2.0, 0.5, 0.05)):
        c = self.c
        s = State(c.M, c.F)
        p, R, v, a, w = self.pose(0.0)
        sp, sv, sth, sbw, sba = sigmas
        sig = np.concatenate([np.full(3, sp), np.full(3, sv), np.full(3, np.deg2rad(sth)), np.full(3, np.deg2rad(sbw)),
                              np.full(3, sba)])
        err = self.rng.normal(0, 1, 15) * sig * c.init_err_scale
        s.x[0:3] = p + err[0:3]
        s.x[3:6] = v + err[3:6]
        dth = err[6:9]
        dR = quat_to_rot(np.array([*(0.5 * dth), 1.0]))
        s.x[6:10] = rot_to_quat(R @ dR)
        s.x[10:13] = self.b_w + err[9:12] * 0.1
        s.x[13:16] = self.b_a + err[12:15]
        s.x[16:20] = self.q_ic
        s.x[20:23] = self.p_ic
        s.x[23:26] = 0.0
        s.x[26:29] = -self.G
        s.x[29] = 0.0
        n = s.n_error_states()
        cov = np.zeros((n, n))
        cov[np.arange(15), np.arange(15)] = sig ** 2
        s.cov = cov
        return s

    # ---- measurements -------------------------------------------------------------------------------------
    def _project(self, lm, frames, noise=True):
        """Normalised observations (len(frames), 2) of landmark lm from the true camera poses."""
        out = np.empty((len(frames), 2))
        for i, k in enumerate(frames):
            pc, Rc = self.cam_pose(self.frame_time(k))
            x = Rc.T @ (lm - pc)
            out[i] = x[:2] / x[2]
        if noise:
            out += self.rng.normal(0, self.c.sigma_img, out.shape)
        return out

    def _sample_visible(self, frames, n):
        """n landmarks on the ground visible from all `frames`."""
        pcs = np.array([self.cam_pose(self.frame_time(k))[0] for k in frames])
        ctr = pcs.mean(axis=0)
        lms = []
        while len(lms) < n:
            cand = np.column_stack([ctr[0] + self.rng.uniform(-7, 7, 4 * n + 16), ctr[1] + self.rng.uniform(-5, 5, 4 * n + 16),
                                    self.rng.uniform(-3.0, 1.0, 4 * n + 16)])
            ok = np.ones(len(cand), dtype=bool)
            for k in (frames[0], frames[len(frames) // 2], frames[-1]):
                pc, Rc = self.cam_pose(self.frame_time(k))
                x = (cand - pc) @ Rc
                ok &= (x[:, 2] > 1.0) & (np.abs(x[:, 0] / x[:, 2]) < 0.9) & (np.abs(x[:, 1] / x[:, 2]) < 0.65)
            lms += list(cand[ok])
        return np.array(lms[:n])

    def _tracks(self, frames, n, outliers=True):
        lms = self._sample_visible(frames, n)
        self._last_lms = lms
        Z = np.empty((n, len(frames), 2))
        for i, k in enumerate(frames):
            pc, Rc = self.cam_pose(self.frame_time(k))
            x = (lms - pc) @ Rc
            Z[:, i, :] = x[:, :2] / x[:, 2:3]
        Z += self.rng.normal(0, self.c.sigma_img, Z.shape)
        if outliers:
            bad = np.nonzero(self.rng.uniform(size=n) < self.c.outlier_frac)[0]
            for j in bad:
                Z[j, self.rng.integers(len(frames))] = self.rng.uniform(-0.8, 0.8, 2)
        return [Z[j] for j in range(n)]

    def measurement(self, k):
        """Track lists for the update at camera frame k (the window then holds frames k-n+1..k)."""
        c = self.c
        n = min(k + 1, c.M)
        frames = list(range(k - n + 1, k + 1))
        m = Measurement(timestamp=self.frame_time(k))
        # landmarks behind the MSCKF / short-MSCKF tracks of the last measurement: lets a multi-agent scenario
        # generate other agents' observations of the same landmarks (MSCKF-MSCKF matches)
        self.last_msckf_lms, self.last_short_lms = np.zeros((0, 3)), []
        if n >= 2 and c.K > 0:
            m.msckf_trks = self._tracks(frames, c.K)
            self.last_msckf_lms = self._last_lms
        if c.n_short > 0 and n >= 3:
            # short tracks are processed BEFORE the window slides / the new clone is added: they end at frame k-1
            nw = min(k, c.M)
            for _ in range(c.n_short):
                L = int(self.rng.integers(2, max(3, nw)))
                L = min(L, nw)
                fr = list(range(k - L, k))
                m.msckf_short_trks += self._tracks(fr, 1, outliers=False)
                self.last_short_lms.append(self._last_lms[0])
        init_frame = c.slam_init_frame if c.slam_init_frame >= 0 else c.M
        if c.F > 0:
            # existing SLAM features: one track per feature slot (last observation is the one used)
            lost = []
            if k > init_frame + 1 and c.churn > 0 and len(self.feat_lm) > c.churn:
                lost = sorted(self.rng.choice(len(self.feat_lm), c.churn, replace=False).tolist())
            for j, lm_id in enumerate(self.feat_lm):
                z = self._project(self.slam_lm[lm_id], [k])[0]
                self.feat_obs[j].append(z)
            keep = [j for j in range(len(self.feat_lm)) if j not in lost]
            m.lost_slam_trk_idxs = lost
            m.slam_trks = [np.array(self.feat_obs[j][-c.M:]) for j in keep]
            self.feat_lm = [self.feat_lm[j] for j in keep]
            self.feat_obs = [self.feat_obs[j] for j in keep]
            n_free = c.F - len(self.feat_lm)
            if k >= init_frame and n_free > 0 and n >= 2:
                n_ms = int(round(n_free * c.slam_msckf_init_frac)) if k == init_frame else 0
                n_std = n_free - n_ms
                new_ms, new_std = [], []
                for i in range(n_free):
                    lm_id = self.next_lm % len(self.slam_lm)
                    self.next_lm += 1
                    z = self._project(self.slam_lm[lm_id], frames)
                    (new_ms if i < n_ms else new_std).append((lm_id, z))
                # feature slots are appended MSCKF-SLAM first, then standard (vio_updater.cpp:425-449)
                for lm_id, z in new_ms + new_std:
                    self.feat_lm.append(lm_id)
                    self.feat_obs.append([z[-1]])
                m.new_msckf_slam_trks = [z for _, z in new_ms]
                m.new_slam_std_trks = [z for _, z in new_std]
                assert n_std == len(new_std)
        t = self.frame_time(k)
        if c.range_every > 0 and k % c.range_every == 0 and len(m.slam_trks) >= 3 and t > 0.1:
            m.range = self._range_measurement(k, len(m.slam_trks))
        if c.sun_every > 0 and k % c.sun_every == 0:
            m.sun_angle = self._sun_measurement(k)
        return m

    # Laser range finder (range_update.cpp:61-139): the beam leaves the camera centre through the normalised image point
    # `pt` and hits the facet spanned by three SLAM landmarks; range = ((f1 - p_c) . n) / (pt . R_c^T n).
    def _range_measurement(self, k, n_tracked):
        c = self.c
        ids = sorted(self.rng.choice(n_tracked, 3, replace=False).tolist())  # slots initialised in earlier frames
        f = [self.slam_lm[self.feat_lm[j]] for j in ids]
        pc, Rc = self.cam_pose(self.frame_time(k))
        z = np.array([(Rc.T @ (fj - pc))[:2] / (Rc.T @ (fj - pc))[2] for fj in f])
        pt = np.array([*z.mean(axis=0), 1.0])
        n = np.cross(f[0] - f[1], f[2] - f[1])
        rng_true = float((f[1] - pc) @ n) / float(pt @ (Rc.T @ n))
        return RangeMeasurement(timestamp=self.frame_time(k), range=rng_true + self.rng.normal(0, c.sigma_range),
                                img_pt_n=(float(pt[0]), float(pt[1])), tr_feat_ids=ids)

    # Sun sensor (solar_update.cpp:44-70): angles of the sun vector in the sensor frame, in degrees; the sensor
    # orientation and the world-frame sun vector are the constants hard-coded there.
    S_Q_I = np.array([-0.063338979194957, 0.007502445522018, 0.930635612981541, 0.360346005598587])  # (x,y,z,w)
    G_SUN = np.array([-0.29385515271891938, -0.55080445540063927, 0.78119370269565391])

    def _sun_measurement(self, k):
        t = self.frame_time(k)
        _, R, *_ = self.pose(t)
        s = quat_to_rot(self.S_Q_I).T @ (R.T @ (self.G_SUN / np.linalg.norm(self.G_SUN)))
        ang = np.degrees([np.arctan2(s[0], s[2]), np.arctan2(s[1], s[2])]) + self.rng.normal(0, self.c.sun_noise_deg, 2)
        return SunAngleMeasurement(timestamp=t, x_angle=float(ang[0]), y_angle=float(ang[1]))

    def imu_between(self, k0, k1, extra=0):
        """IMU samples with timestamps in (t_k0, t_k1 + extra*dt_imu]: list of (t, seq, w_m, a_m)."""
        c = self.c
        i0 = k0 * c.imu_per_frame
        i1 = k1 * c.imu_per_frame + extra
        out = []
        for i in range(i0 + 1, i1 + 1):
            t = i * self.dt_imu
            w_m, a_m = self.imu_sample(t)
            out.append((t, i, w_m, a_m))
        return out


def record(scn: Scenario, n_frames):
    """Generate the reference call sequence (VIO::processImu / processMatchesMeasurement) once:
    a list of ("init", State) / ("imu", t, seq, w_m, a_m) / ("update", Measurement) events.  The update for
    frame k is issued after `latency_imu` further IMU samples have been buffered, so every update
    re-propagates that tail (ekf.cpp:227-255)."""
    c = scn.c
    ev = [("init", scn.initial_state())]
    w0, a0 = scn.imu_sample(0.0)
    ev.append(("imu", 0.0, 0, w0, a0))
    fed = 0
    for k in range(n_frames):
        upto = k * c.imu_per_frame + c.latency_imu
        for i in range(fed + 1, upto + 1):
            t = i * scn.dt_imu
            w_m, a_m = scn.imu_sample(t)
            ev.append(("imu", t, i, w_m, a_m))
        fed = max(fed, upto)
        ev.append(("update", scn.measurement(k)))
    return ev


def replay(events, flt, on_update=None, want_state=True):
    """Feed recorded events to `flt` (x_multi_agent_b200.Filter or any object with the same methods)."""
    k = 0
    for e in events:
        if e[0] == "init":
            flt.initialize_from_state(e[1])
        elif e[0] == "imu":
            flt.process_imu(e[1], e[2], e[3], e[4])
        else:
            flt.set_measurement(e[1])
            st = flt.process_update_measurement()
            if on_update:
                on_update(k, e[1], st)
            k += 1
