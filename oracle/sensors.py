"""Range (laser range finder) and sun-sensor rows of VioUpdater::constructUpdate (oracle; test infrastructure only).

reference: src/x/vio/range_update.cpp:24-265 (RangeUpdate ctor + processRangedFacet), src/x/vio/solar_update.cpp:25-94,
stacked by src/x/vio/vio_updater.cpp:352-403.  Pinned against the reference's own sources compiled in place
(oracle/_ref/libxref.so, tests/test_ref_pinning.py).
"""
from dataclasses import dataclass, field

import numpy as np

from .quat import rot, skew
from .updates import K_CORE, chi2_quantile

K_IDX_Q = 6  # common/types.h:42


@dataclass
class RangeMeasurement:
    """include/x/vio/types.h:223-243 + the facet VioUpdater looks up (vio_updater.cpp:365-366)."""
    timestamp: float = -1.0
    range: float = 0.0
    img_pt_n: tuple = (0.0, 0.0)      # normalised image coordinates of the LRF beam
    tr_feat_ids: list = field(default_factory=list)  # TrackManager::featureTriangleAtPoint: 3 SLAM feature ids, or empty


@dataclass
class SunAngleMeasurement:
    """include/x/vio/types.h:250-254."""
    timestamp: float = -1.0
    x_angle: float = 0.0
    y_angle: float = 0.0


def _mat_ivd(a, b, r):
    m = np.eye(3)
    m[0, 2] = -a / r
    m[1, 2] = -b / r
    m[2, 2] = -1.0 / r
    return m


class RangeUpdate:
    """reference: range_update.cpp:24-265 (processRangedFacet): one row, gated with chi2(0.9, 1)."""

    def __init__(self, meas, quats, poss, feature_states, anchor_idxs, P, n_poses_max, sigma_range):
        cols = P.shape[1]
        self.jac = np.zeros((1, cols))
        self.cov_m_diag = np.ones(1)
        self.res = np.zeros(1)
        ids = list(meas.tr_feat_ids)
        G_p_fj, R_a, anchor_idx, alpha, beta, rho = [], [], [], [], [], []
        for j in range(3):  # :76-97
            a, b, r = feature_states[3 * ids[j]:3 * ids[j] + 3]
            alpha.append(a), beta.append(b), rho.append(r)
            ai = anchor_idxs[ids[j]]
            anchor_idx.append(ai)
            R_a.append(rot(quats[ai]))
            G_p_fj.append(1.0 / r * R_a[j] @ np.array([a, b, 1.0]) + poss[ai])
        R_i = rot(quats[-1])  # :103-110
        G_p_Ci = np.asarray(poss[-1], dtype=float)
        G_n = np.cross(G_p_fj[0] - G_p_fj[1], G_p_fj[2] - G_p_fj[1])  # :123
        pt = np.array([meas.img_pt_n[0], meas.img_pt_n[1], 1.0])
        a_ = float((G_p_fj[1] - G_p_Ci) @ G_n)  # :129-131
        b_ = float(pt @ (R_i.T @ G_n))
        range_hat = a_ / b_
        res_j = meas.range - range_hat
        h_j = np.zeros((1, cols))
        J_pc = -1.0 / b_ * G_n  # :148
        J_qc = a_ / b_ ** 2 * G_n @ R_i @ skew(pt)  # :151-153
        G_p_r = a_ / b_ * R_i @ pt + G_p_Ci
        bary = 1.0 / 3.0 * (G_p_fj[0] + G_p_fj[1] + G_p_fj[2])
        edges = [(2, 1), (0, 2), (1, 0)]  # :162, :178, :194
        pos = len(quats) - 1
        c = K_CORE + pos * 3  # :209-215
        h_j[0, c:c + 3] = J_pc
        c += n_poses_max * 3
        h_j[0, c:c + 3] = J_qc
        for j in range(3):
            p, q = edges[j]
            J_f = 1.0 / b_ * (1.0 / 3.0 * G_n + np.cross(G_p_fj[p] - G_p_fj[q], bary - G_p_r))
            J_qa = -1.0 / rho[j] * J_f @ R_a[j] @ skew(np.array([alpha[j], beta[j], 1.0]))
            J_fi = 1.0 / rho[j] * J_f @ R_a[j] @ _mat_ivd(alpha[j], beta[j], rho[j])
            c = K_CORE + anchor_idx[j] * 3  # :217-224 (anchor blocks accumulate, the feature block is assigned)
            h_j[0, c:c + 3] += J_f
            c += n_poses_max * 3
            h_j[0, c:c + 3] += J_qa
            c = K_CORE + (n_poses_max * 2 + ids[j]) * 3
            h_j[0, c:c + 3] = J_fi
        var_range = sigma_range * sigma_range  # :246-250
        S = float((h_j @ P @ h_j.T)[0, 0]) + var_range
        self.gamma = res_j * (1.0 / S) * res_j
        self.chi = chi2_quantile(0.9, 1)
        self.inlier = bool(self.gamma < self.chi)
        if self.inlier:
            self.jac = h_j
            self.res = np.array([res_j])
            self.cov_m_diag = np.array([var_range])


class SolarUpdate:
    """reference: solar_update.cpp:25-94: two rows on the core attitude, no gate; calibration constants as hard-coded."""

    VAR_SUN_ANGLE = 10000 * 0.01777777777  # :48
    S_Q_I = np.array([-0.063338979194957, 0.007502445522018, 0.930635612981541, 0.360346005598587])  # (x,y,z,w), :51
    G_SUN = np.array([-0.29385515271891938, -0.55080445540063927, 0.78119370269565391])  # :55
    RAD2DEG = 57.2957795130  # :67

    def __init__(self, angle, quat, n_cols):
        g_sun = self.G_SUN / np.sqrt(self.G_SUN @ self.G_SUN)
        R_s = rot(self.S_Q_I)
        R_q = rot(quat)
        s = R_s.T @ (R_q.T @ g_sun)
        s = s / np.sqrt(s @ s)
        angles_hat = np.array([self.RAD2DEG * np.arctan2(s[0], s[2]), self.RAD2DEG * np.arctan2(s[1], s[2])])
        self.res = np.array([angle.x_angle, angle.y_angle]) - angles_hat
        mat = np.zeros((2, 3))
        mat[0, 0] = s[2] / (s[0] ** 2 + s[2] ** 2)
        mat[1, 1] = s[2] / (s[1] ** 2 + s[2] ** 2)
        mat[0, 2] = -s[0] / (s[0] ** 2 + s[2] ** 2)
        mat[1, 2] = -s[1] / (s[1] ** 2 + s[2] ** 2)
        sv = R_q.T @ g_sun
        J_att = self.RAD2DEG * mat @ R_s.T @ skew(sv)
        self.jac = np.zeros((2, n_cols))
        self.jac[:, K_IDX_Q:K_IDX_Q + 3] = J_att
        self.cov_m_diag = self.VAR_SUN_ANGLE * np.ones(2)
