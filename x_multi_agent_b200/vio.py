"""Host-side mirror of the x::VIO facade for callers that deliver feature MATCHES.

reference: include/x/vio/vio.h, src/x/vio/vio.cpp:40-215 (setUp, initAtTime), :274-323 (processMatchesMeasurement), :343-370
(processImu), :372-434 (importMatches), :576-707 (loadParamsFromYaml); src/x/vio/vio_updater.cpp:126-179 (preProcess: camera
attitude list -> TrackManager::manageTracks -> the five track lists).

What runs where: parameters, the camera attitude list and the track management are host code (TrackManager: xb_tm_* in
libxb200.so); everything from the track lists on is the device filter (Filter: xb_ekf_* / xb_vio_*).  Out of scope here, as
in SURVEY 2: images (Tracker / KLT), place recognition.  Range and sun-sensor measurements (setLastRangeMeasurement /
setLastSunAngleMeasurement, vio.cpp:217-224) are forwarded to the filter.
"""
import numpy as np

from .filter import Filter, RangeMeasurement, State, SunAngleMeasurement
from .track_manager import TrackManager

# keys of the reference's YAML files (vio.cpp:576-707) -> (default, length); vectors are lists, quaternions are [w, x, y, z]
_VEC = {"p": 3, "v": 3, "q": 4, "b_w": 3, "b_a": 3, "sigma_dp": 3, "sigma_dv": 3, "sigma_dtheta": 3, "sigma_dbw": 3, "sigma_dba": 3,
        "cam1_p_ic": 3, "cam1_q_ic": 4, "g": 3}
_DEFAULTS = {"p": [0, 0, 0], "v": [0, 0, 0], "q": [1, 0, 0, 0], "b_w": [0, 0, 0], "b_a": [0, 0, 0], "sigma_dp": [0, 0, 0],
             "sigma_dv": [0.05, 0.05, 0.05], "sigma_dtheta": [3, 3, 3], "sigma_dbw": [6, 6, 6], "sigma_dba": [0.3, 0.3, 0.3],
             "cam1_fx": 0.46, "cam1_fy": 0.61, "cam1_cx": 0.5, "cam1_cy": 0.5, "cam1_s": 0.0, "cam1_img_height": 480,
             "cam1_img_width": 640, "cam1_p_ic": [0, 0, 0], "cam1_q_ic": [1, 0, 0, 0], "cam1_time_offset": 0.0, "sigma_img": 0.02,
             "n_a": 0.0083, "n_ba": 0.00083, "n_w": 0.0013, "n_bw": 0.00013, "n_tiles_h": 1, "n_tiles_w": 1, "msckf_baseline": 30.0,
             "min_track_length": 10, "rho_0": 0.5, "sigma_rho_0": 0.5, "iekf_iter": 1, "n_poses_max": 10, "n_slam_features_max": 15,
             "g": [0, 0, -9.81], "state_buffer_size": 250, "sigma_range": 0.05, "max_feat_per_tile": 40}


def load_params_from_yaml(path):
    """VIO::loadParamsFromYaml (vio.cpp:576-707) without cv::FileStorage: a plain YAML mapping with the reference's keys
    (the `%YAML:1.0` directive line OpenCV writes is skipped); missing keys keep the defaults above."""
    import yaml
    text = "\n".join(line for line in open(path).read().splitlines() if not line.startswith("%YAML"))
    doc = yaml.safe_load(text) or {}
    params = dict(_DEFAULTS)
    params.update(doc)
    for k, n in _VEC.items():
        if len(params[k]) != n:
            raise ValueError(f"parameter '{k}' needs {n} values")
    return params


def _qmul(a, b):   # (x, y, z, w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


class VIO:
    def __init__(self):
        self.params = None
        self.filter = None
        self.track_manager = None
        self.initialized = False

    def set_up(self, params, device=0, **filter_kw):
        """VIO::setUp (vio.cpp:113-215)."""
        p = dict(_DEFAULTS)
        p.update(params)
        if p["min_track_length"] > p["n_poses_max"]:
            raise ValueError("'min_track_length' cannot be larger than 'n_poses_max'")   # vio.cpp:124-127
        self.params = p
        # minimum MSCKF baseline in the normal plane (vio.cpp:163-168)
        bx = p["msckf_baseline"] / (p["cam1_img_width"] * p["cam1_fx"])
        by = p["msckf_baseline"] / (p["cam1_img_height"] * p["cam1_fy"])
        if self.track_manager:
            self.track_manager.close()
        self.track_manager = TrackManager(p["cam1_fx"], p["cam1_fy"], p["cam1_cx"], p["cam1_cy"], p["cam1_s"], p["cam1_img_width"],
                                          p["cam1_img_height"], bx, by, p["n_tiles_h"], p["n_tiles_w"])
        if self.filter:
            self.filter.close()
        self.filter = Filter(p["n_poses_max"], p["n_slam_features_max"], n_slots=p["state_buffer_size"], device=device,
                             sigma_img=p["sigma_img"], sigma_range=p["sigma_range"], rho_0=p["rho_0"], sigma_rho_0=p["sigma_rho_0"], iekf_iter=p["iekf_iter"],
                             min_track_length=p["min_track_length"],
                             a_m_max=50.0, delta_seq_imu=1, time_margin=0.02,     # Ekf::set arguments of vio.cpp:208-214
                             n_w=p["n_w"], n_bw=p["n_bw"], n_a=p["n_a"], n_ba=p["n_ba"], g=tuple(p["g"]), **filter_kw)
        self.initialized = False

    def initial_state(self, time):
        """The state VIO::initAtTime builds (vio.cpp:63-96): zero vision states, diagonal initial covariance, gravity-aligned
        first IMU sample."""
        p = self.params
        M, F = p["n_poses_max"], p["n_slam_features_max"]
        s = State(M, F)
        s.x[0:3], s.x[3:6] = p["p"], p["v"]
        q = np.array([p["q"][1], p["q"][2], p["q"][3], p["q"][0]], dtype=float)       # YAML order is w, x, y, z
        s.x[6:10] = q / np.linalg.norm(q)
        s.x[10:13], s.x[13:16] = p["b_w"], p["b_a"]
        qic = np.array([p["cam1_q_ic"][1], p["cam1_q_ic"][2], p["cam1_q_ic"][3], p["cam1_q_ic"][0]], dtype=float)
        s.x[16:20] = qic / np.linalg.norm(qic)
        s.x[20:23] = p["cam1_p_ic"]
        s.x[23:26] = 0.0
        s.x[26:29] = -np.asarray(p["g"], dtype=float)
        s.x[29], s.x[30] = time, 0
        sig = np.concatenate([p["sigma_dp"], p["sigma_dv"], np.asarray(p["sigma_dtheta"]) * np.pi / 180.0,
                              np.asarray(p["sigma_dbw"]) * np.pi / 180.0, p["sigma_dba"], np.zeros(6 * M + 3 * F)])
        s.cov = np.diag(sig * sig)
        return s

    def init_at_time(self, time):
        """VIO::initAtTime (vio.cpp:54-111)."""
        self.track_manager.clear()
        self.filter.initialize_from_state(self.initial_state(time))
        self.initialized = True

    def process_imu(self, timestamp, seq, w_m, a_m):
        """VIO::processImu (vio.cpp:343-370)."""
        if not self.initialized:
            return None
        return self.filter.process_imu(timestamp, seq, w_m, a_m)

    def camera_attitudes(self, state, n_poses):
        """The list VioUpdater::preProcess builds (vio_updater.cpp:142-153): the window's camera attitudes cropped to
        n_poses_max - 1 entries from the end (convertCameraAttitudesToList, state_manager.cpp:540-565), then the current one
        (State::computeCameraAttitude, state.cpp:184-187)."""
        M = self.params["n_poses_max"]
        qa = state.q_array.reshape(M, 4)[:n_poses]
        size_out = min(M - 1, n_poses) if M - 1 > 0 else n_poses
        q = state.q / np.linalg.norm(state.q)
        qic = state.q_ic / np.linalg.norm(state.q_ic)
        return np.vstack([qa[n_poses - size_out:], _qmul(q, qic)[None]])

    def set_last_range_measurement(self, timestamp, range_m):
        """VIO::setLastRangeMeasurement (vio.cpp:217-219): kept until replaced, applied with every image after it."""
        self.last_range = (float(timestamp), float(range_m))

    def set_last_sun_angle_measurement(self, timestamp, x_angle, y_angle):
        """VIO::setLastSunAngleMeasurement (vio.cpp:221-224)."""
        self.last_sun = SunAngleMeasurement(float(timestamp), float(x_angle), float(y_angle))

    def _sensors(self, m):
        """The range / sun-angle members of the VioMeasurement (vio.cpp:288-298) and the facet lookup of
        VioUpdater::constructUpdate (vio_updater.cpp:358-369: the hard-coded image point (320.5, 240.5))."""
        p, tm = self.params, self.track_manager
        rng = getattr(self, "last_range", None)
        if rng is not None and rng[0] > 0.1 and m.slam_trks:
            ids = tm.feature_triangle_at_point(320.5, 240.5)
            if ids:
                pt = tm.normalize_point((p["cam1_img_width"] + 1) / 2.0, (p["cam1_img_height"] + 1) / 2.0)
                m.range = RangeMeasurement(rng[0], rng[1], pt, ids)
        sun = getattr(self, "last_sun", None)
        if sun is not None and sun.timestamp > -1:
            m.sun_angle = sun
        return m

    def process_matches_measurement(self, timestamp, seq, match_vector):
        """VIO::processMatchesMeasurement (vio.cpp:274-323): time correction, match import (skipped for the first image, before
        any pose is in the window), track management, Ekf::processUpdateMeasurement."""
        if not self.initialized:
            return None
        p = self.params
        t = timestamp + p["cam1_time_offset"]
        flt, tm = self.filter, self.track_manager
        mv = np.asarray(match_vector, dtype=float).reshape(-1, 10)
        if flt.n_poses == 0:
            mv = mv[:0]
        state = flt.update_begin(t)                    # the buffered state closest to the image (ekf.cpp:183-199)
        if state is None:
            return None
        rots = self.camera_attitudes(state, flt.n_poses)
        tm.manage_tracks(mv, rots, p["n_poses_max"], p["n_slam_features_max"], p["min_track_length"])
        flt.set_measurement(self._sensors(tm.measurement(t, p["n_poses_max"])))
        flt.updater_update()                           # Updater::update (updater.cpp:39-115)
        updated = flt.update_end()
        if updated is not None:
            updated.time = timestamp                   # vio.cpp:314-316
        return updated

    def close(self):
        if self.filter:
            self.filter.close()
        if self.track_manager:
            self.track_manager.close()
        self.filter = self.track_manager = None
