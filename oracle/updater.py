"""Kalman update + the VioUpdater template method (oracle; test infrastructure only).

reference: src/x/ekf/updater.cpp, src/x/vio/vio_updater.cpp:200-512.
The front end (tracker / track manager) is out of scope: the five track lists + lost-feature
indexes that `VioUpdater::preProcess` leaves behind (vio_updater.cpp:172-179) are injected directly.
"""
from dataclasses import dataclass, field

import numpy as np

from .sensors import RangeMeasurement, RangeUpdate, SolarUpdate, SunAngleMeasurement
from .state_manager import StateManager
from .updates import MsckfSlamUpdate, MsckfUpdate, SlamUpdate


def apply_qr_decomposition(h, res, R_diag, sigma_img):
    """reference: vio_updater.cpp:487-512.  Returns (h, res, R) -- R as a dense matrix like the reference."""
    rows, cols = h.shape
    if rows > cols + 1:
        hres = np.hstack([h, res[:, None]])
        thz = np.triu(np.linalg.qr(hres, mode="r"))
        h = thz[:cols, :cols].copy()
        res = thz[:cols, cols].copy()
        R = sigma_img * sigma_img * np.eye(cols)
    else:
        R = np.diag(R_diag)
    return h, res, R


def apply_update(state, H, res, R, correction_total, cov_update=True):
    """reference: updater.cpp:117-141 (explicit inverse, dense (I-KH)P, symmetrise)."""
    P = state.cov
    S = H @ P @ H.T + R
    K = P @ H.T @ np.linalg.inv(S)
    correction = K @ (res + H @ correction_total) - correction_total
    n = P.shape[0]
    if cov_update:
        P = (np.eye(n) - K @ H) @ P
        P = 0.5 * (P + P.T)
        state.cov = P
    state.correct(correction)
    correction_total += correction
    return correction


def apply_ci(state, ci_P, H, res, S):
    """reference: updater.cpp:144-161."""
    K = ci_P @ H.T @ np.linalg.inv(S)
    correction = K @ res
    n = ci_P.shape[0]
    P = (np.eye(n) - K @ H) @ ci_P
    state.cov = 0.5 * (P + P.T)
    state.correct(correction)
    return correction


@dataclass
class VisualMeasurement:
    """Output of VioUpdater::preProcess (vio_updater.cpp:172-179), injected at the seam."""
    timestamp: float = 0.0
    slam_trks: list = field(default_factory=list)            # one per active SLAM feature (index = feature id)
    msckf_trks: list = field(default_factory=list)
    msckf_short_trks: list = field(default_factory=list)
    new_slam_std_trks: list = field(default_factory=list)
    new_msckf_slam_trks: list = field(default_factory=list)
    lost_slam_trk_idxs: list = field(default_factory=list)
    range: RangeMeasurement = field(default_factory=RangeMeasurement)      # VioMeasurement::range (vio/types.h:300)
    sun_angle: SunAngleMeasurement = field(default_factory=SunAngleMeasurement)  # VioMeasurement::sun_angle (:305)


class VioUpdaterOracle:
    """Restates Updater::update (updater.cpp:39-115, single-UAV build) with VioUpdater's overrides."""

    def __init__(self, n_poses_max, n_features_max, sigma_img, rho_0=0.5, sigma_rho_0=0.25, iekf_iter=1,
                 ci_msckf_w=0.1, sigma_range=0.05):
        self.ci_msckf_w = ci_msckf_w
        self.sigma_range = sigma_range
        self.msckf_matches = []  # VioUpdater::msckf_matches_ (vio_updater.h:282), consumed by the constructors
        self.sm = StateManager(n_poses_max, n_features_max)
        self.sigma_img = sigma_img
        self.rho_0 = rho_0
        self.sigma_rho_0 = sigma_rho_0
        self.iekf_iter = iekf_iter
        self.meas = VisualMeasurement()
        self.last = {}

    def set_measurement(self, meas):
        self.meas = meas

    def get_time(self):
        return self.meas.timestamp

    # vio_updater.cpp:217-264 (single-UAV)
    def construct_short_msckf_update(self, state):
        quats = self.sm.camera_attitudes(state)
        poss = self.sm.camera_positions(state)
        msckf = MsckfUpdate(self.meas.msckf_short_trks, quats, poss, state.cov, state.n_poses_max(), self.sigma_img)
        self.last["short"] = msckf
        return apply_qr_decomposition(msckf.jac, msckf.res, msckf.cov_m_diag, self.sigma_img)

    # vio_updater.cpp:352-403: the range row (one, gated) and the two sun-sensor rows; each measurement is used once
    def _sensor_rows(self, state, quats, poss):
        P = state.cov
        cols = P.shape[1]
        h_l, res_l, r_l = np.zeros((0, cols)), np.zeros(0), np.zeros(0)
        rng, sun = self.meas.range, self.meas.sun_angle
        if rng.timestamp > 0.1 and self.meas.slam_trks and len(rng.tr_feat_ids) > 0:
            ru = RangeUpdate(rng, quats, poss, state.f_array, self.sm.anchor_idxs, P, state.n_poses_max(), self.sigma_range)
            self.last["range"] = ru
            h_l, res_l, r_l = ru.jac, ru.res, ru.cov_m_diag
            rng.timestamp = -1.0
        h_s, res_s, r_s = np.zeros((0, cols)), np.zeros(0), np.zeros(0)
        if sun.timestamp > -1:
            su = SolarUpdate(sun, state.q, cols)
            self.last["solar"] = su
            h_s, res_s, r_s = su.jac, su.res, su.cov_m_diag
            sun.timestamp = -1.0
        return np.vstack([h_l, h_s]), np.concatenate([res_l, res_s]), np.concatenate([r_l, r_s])

    # vio_updater.cpp:266-423
    def construct_update(self, state):
        quats = self.sm.camera_attitudes(state)
        poss = self.sm.camera_positions(state)
        P = state.cov
        M = state.n_poses_max()
        msckf = MsckfUpdate(self.meas.msckf_trks, quats, poss, P, M, self.sigma_img)
        msckf_slam = MsckfSlamUpdate(self.meas.new_msckf_slam_trks, quats, poss, P, M, self.sigma_img)
        slam = SlamUpdate(self.meas.slam_trks, quats, poss, state.f_array, self.sm.anchor_idxs, P, M, self.sigma_img)
        self.last.update(msckf=msckf, msckf_slam=msckf_slam, slam=slam)
        h_x, res_x, r_x = self._sensor_rows(state, quats, poss)
        h = np.vstack([msckf.jac, msckf_slam.jac, slam.jac, h_x])
        r_diag = np.concatenate([msckf.cov_m_diag, msckf_slam.cov_m_diag, slam.cov_m_diag, r_x])
        res = np.concatenate([msckf.res, msckf_slam.res, slam.res, res_x])
        return apply_qr_decomposition(h, res, r_diag, self.sigma_img)

    # ---- MULTI_UAV build: vio_updater.cpp:217-264 / 266-423 with the four CI lists -------------------------
    def _construct_multi(self, state, which):
        from .ci import MultiMsckfUpdate
        quats = self.sm.camera_attitudes(state)
        poss = self.sm.camera_positions(state)
        P = state.cov
        M = state.n_poses_max()
        trks = self.meas.msckf_short_trks if which == 1 else self.meas.msckf_trks
        ids = [(which, j) for j in range(len(trks))]
        msckf = MultiMsckfUpdate(trks, ids, quats, poss, P, M, self.sigma_img, self.msckf_matches, self.ci_msckf_w)
        lists = (msckf.S_list, msckf.P_list, msckf.H_list, msckf.res_list)
        if which == 1:
            self.last["short"] = msckf
            return apply_qr_decomposition(msckf.jac, msckf.res, msckf.cov_m_diag, self.sigma_img), lists
        msckf_slam = MsckfSlamUpdate(self.meas.new_msckf_slam_trks, quats, poss, P, M, self.sigma_img)
        slam = SlamUpdate(self.meas.slam_trks, quats, poss, state.f_array, self.sm.anchor_idxs, P, M, self.sigma_img)
        self.last.update(msckf=msckf, msckf_slam=msckf_slam, slam=slam)
        h_x, res_x, r_x = self._sensor_rows(state, quats, poss)
        h = np.vstack([msckf.jac, msckf_slam.jac, slam.jac, h_x])
        r_diag = np.concatenate([msckf.cov_m_diag, msckf_slam.cov_m_diag, slam.cov_m_diag, r_x])
        res = np.concatenate([msckf.res, msckf_slam.res, slam.res, res_x])
        return apply_qr_decomposition(h, res, r_diag, self.sigma_img), lists

    def update_multi_uav(self, state):
        """Updater::update as compiled with -DMULTI_UAV (updater.cpp:39-115): the short-MSCKF step applies ONLY the
        CI lists (its stacked h is built and dropped, :58-70); the main step applies the CI lists first, then ONE
        applyUpdate with the (h, res) linearised BEFORE the CI corrections (:84-97); no IEKF loop."""
        self.last.pop("range", None)
        self.last.pop("solar", None)
        correction = np.zeros(state.n_error_states())
        if self.meas.msckf_short_trks:
            _, (S_l, P_l, H_l, r_l) = self._construct_multi(state, 1)
            for S_j, P_j, h_j, r_j in zip(S_l, P_l, H_l, r_l):
                apply_ci(state, P_j, h_j, r_j, S_j)
        self.sm.manage(state, list(self.meas.lost_slam_trk_idxs))
        m = self.meas
        if m.msckf_trks or m.slam_trks or m.new_slam_std_trks or m.new_msckf_slam_trks:
            correction = np.zeros(state.n_error_states())
            (h, res, r), (S_l, P_l, H_l, r_l) = self._construct_multi(state, 0)
            self.last.update(h=h, res=res, r=r, ci_lists=(S_l, P_l, H_l, r_l))
            for S_j, P_j, h_j, r_j in zip(S_l, P_l, H_l, r_l):
                apply_ci(state, P_j, h_j, r_j, S_j)
            if h.size > 0:
                apply_update(state, h, res, r, correction, True)
            self.post_update(state, correction)
        self.last["correction"] = correction
        return state

    # vio_updater.cpp:425-449
    def post_update(self, state, correction):
        if self.meas.new_msckf_slam_trks:
            ms = self.last["msckf_slam"]
            self.sm.init_msckf_slam_features(state, ms.H1, ms.H2, ms.r1, ms.features, correction, self.sigma_img)
        if self.meas.new_slam_std_trks:
            ivds = SlamUpdate.compute_inverse_depths_new(self.meas.new_slam_std_trks, self.rho_0)
            self.sm.init_standard_slam_features(state, ivds, self.sigma_img, self.sigma_rho_0)

    # updater.cpp:39-115
    def update(self, state):
        self.last.pop("range", None)
        self.last.pop("solar", None)
        correction = np.zeros(state.n_error_states())
        if self.meas.msckf_short_trks:  # preUpdateShortMsckf, vio_updater.cpp:209-215
            h, res, r = self.construct_short_msckf_update(state)
            if h.size > 0:
                apply_update(state, h, res, r, correction, True)
        self.sm.manage(state, list(self.meas.lost_slam_trk_idxs))  # preUpdate, vio_updater.cpp:200-207
        m = self.meas
        requested = bool(m.msckf_trks or m.slam_trks or m.new_slam_std_trks or m.new_msckf_slam_trks)
        if requested:
            correction = np.zeros(state.n_error_states())
            for i in range(self.iekf_iter):
                h, res, r = self.construct_update(state)
                self.last.update(h=h, res=res, r=r)
                if h.size > 0:
                    apply_update(state, h, res, r, correction, i == self.iekf_iter - 1)
            self.post_update(state, correction)
        self.last["correction"] = correction
        return state
