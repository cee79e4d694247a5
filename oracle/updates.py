"""Per-measurement H / r builders (oracle; test infrastructure only).

reference: src/x/vio/msckf_update.cpp, src/x/vio/slam_update.cpp, src/x/vio/msckf_slam_update.cpp.
Tracks are (L,2) float arrays of normalised image coordinates, oldest observation first.
Pose lists hold the `n_poses` active window poses (state_manager.cpp:539-584), quaternions (x,y,z,w).
"""
import numpy as np
from scipy.stats import chi2 as _chi2

from .quat import rot, skew
from .state import K_CORE
from .state_manager import _mat_ivd
from .triangulation import Triangulation

G_OC = np.array([0.0, 0.0, -9.81])  # hard-coded in the OC projection, msckf_update.cpp:393
# Mirror of xb_config.oc_projection (include/xb200.h).  True = the reference as written.  False is NOT a reference
# option: it exists so that the benchmark can run the path on a consistent filter; the device has the
# same switch and the two are always set together.
OC_PROJECTION = True


def chi2_quantile(p, dof):
    """boost::math::quantile(chi_squared_distribution<>(dof), p) (third-party: Boost >= 1.71, unpinned,
    CMakeLists.txt:117; call sites msckf_update.cpp:459-461, slam_update.cpp:196-197)."""
    return float(_chi2.ppf(p, dof))


def _vis_jac(cp):
    """J_i of eq. 22/23 (msckf_update.cpp:365-373)."""
    return np.array([[1.0 / cp[2], 0.0, -cp[0] / cp[2] ** 2],
                     [0.0, 1.0 / cp[2], -cp[1] / cp[2] ** 2]])


def global_feature_position(ivd, q_cn, p_cn):
    """reference: msckf_update.cpp:283-304."""
    a, b, r = ivd
    return 1.0 / r * rot(q_cn) @ np.array([a, b, 1.0]) + p_cn


def msckf_track_jacobians(track, quats, poss, n_poses_max, n_cols, G_p_fj):
    """Observation loop of MsckfUpdate::processOneTrack (msckf_update.cpp:328-417).

    Returns (jac_j 2L x N, Hf_j 2L x 3, res_j 2L) or None when the camera-frame point has a NaN (:349-358)."""
    L = track.shape[0]
    n_p = len(poss)
    jac = np.zeros((2 * L, n_cols))
    Hf = np.zeros((2 * L, 3))
    res = np.zeros(2 * L)
    for i in range(L):
        pos = n_p - L + i
        R = rot(quats[pos])
        p_c = poss[pos]
        cp = R.T @ (G_p_fj - p_c)
        if np.any(np.isnan(cp)):
            return None
        res[2 * i] = track[i, 0] - cp[0] / cp[2]
        res[2 * i + 1] = track[i, 1] - cp[1] / cp[2]
        J_i = _vis_jac(cp)
        J_pos = -J_i @ R.T
        J_att = J_i @ skew(cp)
        # Observability-constrained projection (Hesch 2012), msckf_update.cpp:393-406
        if OC_PROJECTION:
            u_pos = (R @ G_OC)[:, None]
            J_pos = J_pos - J_pos @ u_pos @ np.linalg.inv(u_pos.T @ u_pos) @ u_pos.T
            u_att = (skew(G_p_fj - p_c) @ G_OC)[:, None]
            J_att = J_att - J_att @ u_att @ np.linalg.inv(u_att.T @ u_att) @ u_att.T
        Hf[2 * i:2 * i + 2] = -J_pos
        c = K_CORE + pos * 3
        jac[2 * i:2 * i + 2, c:c + 3] = J_pos
        c += n_poses_max * 3
        jac[2 * i:2 * i + 2, c:c + 3] = J_att
    return jac, Hf, res


def householder_q(Hf):
    """`Hf.householderQr().householderQ()` as a full m x m matrix (msckf_update.cpp:423).  Basis and
    signs are implementation-defined (Eigen vs LAPACK); every consumer is invariant to that choice."""
    q, _ = np.linalg.qr(Hf, mode="complete")
    return q


class MsckfUpdate:
    """reference: msckf_update.cpp:27-63 (single-agent build).  Dense formulation on purpose."""

    def __init__(self, trks, quats, poss, P, n_poses_max, sigma_img, max_iter=10, term=1e-5):
        n_trks = len(trks)
        n_obs = sum(t.shape[0] for t in trks)
        rows = 2 * n_obs - 3 * n_trks
        cols = P.shape[1]
        self.jac = np.zeros((rows, cols))
        self.cov_m_diag = np.ones(rows)
        self.res = np.zeros(rows)
        self.inlier = np.zeros(n_trks, dtype=bool)
        self.gamma = np.full(n_trks, np.nan)
        self.chi = np.full(n_trks, np.nan)
        self.features_ivd = np.full((n_trks, 3), np.nan)
        self.G_p_f = np.full((n_trks, 3), np.nan)
        var_img = sigma_img * sigma_img
        row_h = 0
        for j, trk in enumerate(trks):
            trk = np.asarray(trk, dtype=float)
            L = trk.shape[0]
            # preProcessOneTrack, msckf_update.cpp:145-166: fresh triangulator over the last L poses
            tri = Triangulation(quats[len(quats) - L:], poss[len(poss) - L:], max_iter, term)
            ivd = tri.triangulate_gn(trk)
            G_p_fj = global_feature_position(ivd, quats[-1], poss[-1])
            self.features_ivd[j], self.G_p_f[j] = ivd, G_p_fj
            out = msckf_track_jacobians(trk, quats, poss, n_poses_max, cols, G_p_fj)
            if out is None:
                continue
            jac_j, Hf_j, res_j = out
            q = householder_q(Hf_j)
            A = q[:, 3:]
            res0 = A.T @ res_j
            jac0 = A.T @ jac_j
            S = jac0 @ P @ jac0.T + var_img * np.eye(2 * L - 3)  # dense N^2 product, msckf_update.cpp:457
            gamma = float(res0 @ np.linalg.inv(S) @ res0)
            chi = chi2_quantile(0.95, 2.0 * L - 3.0)
            self.gamma[j], self.chi[j] = gamma, chi
            if gamma < chi:
                self.inlier[j] = True
                self.jac[row_h:row_h + 2 * L - 3] = jac0
                self.res[row_h:row_h + 2 * L - 3] = res0
                self.cov_m_diag[row_h:row_h + 2 * L - 3] = var_img
                row_h += 2 * L - 3
        self.rows_used = row_h


class SlamUpdate:
    """reference: slam_update.cpp:25-214.  `trks[j]` belongs to SLAM feature j; only the last observation is used."""

    def __init__(self, trks, quats, poss, feature_states, anchor_idxs, P, n_poses_max, sigma_img):
        n_trks = len(trks)
        cols = P.shape[1]
        self.jac = np.zeros((2 * n_trks, cols))
        self.cov_m_diag = np.ones(2 * n_trks)
        self.res = np.zeros(2 * n_trks)
        self.inlier = np.zeros(n_trks, dtype=bool)
        self.gamma = np.full(n_trks, np.nan)
        self.chi = np.full(n_trks, np.nan)
        var_img = sigma_img * sigma_img
        row_h = 0
        for j, trk in enumerate(trks):
            trk = np.asarray(trk, dtype=float)
            h_j = np.zeros((2, cols))
            a, b, r = feature_states[3 * j:3 * j + 3]
            anchor = anchor_idxs[j]
            R_a = rot(quats[anchor])
            G_p_fj = 1.0 / r * R_a @ np.array([a, b, 1.0]) + poss[anchor]
            R_n = rot(quats[-1])
            cp = R_n.T @ (G_p_fj - poss[-1])
            track_size = trk.shape[0]
            res_j = np.array([trk[-1, 0] - cp[0] / cp[2], trk[-1, 1] - cp[1] / cp[2]])
            pos = len(quats) - 1
            fcol = K_CORE + (n_poses_max * 2 + j) * 3
            if anchor == pos:  # slam_update.cpp:120-130
                h_j[0, fcol] = 1.0
                h_j[1, fcol + 1] = 1.0
            else:
                J_i = _vis_jac(cp)
                J_att = J_i @ skew(cp)
                J_pos = -J_i @ R_n.T
                J_anchor_att = -1.0 / r * J_i @ R_n.T @ R_a @ skew(np.array([a, b, 1.0]))
                J_anchor_pos = -J_pos
                Hf = 1.0 / r * J_i @ R_n.T @ R_a @ _mat_ivd(a, b, r)
                c = K_CORE + pos * 3
                h_j[:, c:c + 3] = J_pos
                c += n_poses_max * 3
                h_j[:, c:c + 3] = J_att
                c = K_CORE + anchor * 3
                h_j[:, c:c + 3] = J_anchor_pos
                c += n_poses_max * 3
                h_j[:, c:c + 3] = J_anchor_att
                h_j[:, fcol:fcol + 3] = Hf
            S = h_j @ P @ h_j.T + var_img * np.eye(2)
            gamma = float(res_j @ np.linalg.inv(S) @ res_j)
            chi = chi2_quantile(0.9, 2 * track_size)  # slam_update.cpp:196-197 (0.9, 2*track_size dof)
            self.gamma[j], self.chi[j] = gamma, chi
            if gamma < chi:
                self.inlier[j] = True
                self.jac[row_h:row_h + 2] = h_j
                self.res[row_h:row_h + 2] = res_j
                self.cov_m_diag[row_h:row_h + 2] = var_img
                row_h += 2
        self.rows_used = row_h

    @staticmethod
    def compute_inverse_depths_new(new_trks, rho_0):
        """reference: slam_update.cpp:216-242."""
        ivds = np.zeros(3 * len(new_trks))
        for j, trk in enumerate(new_trks):
            ivds[3 * j:3 * j + 3] = (trk[-1][0], trk[-1][1], rho_0)
        return ivds


class MsckfSlamUpdate:
    """reference: msckf_slam_update.cpp:26-267 (Li 2012 feature promotion)."""

    def __init__(self, trks, quats, poss, P, n_poses_max, sigma_img, max_iter=10, term=1e-5):
        n_trks = len(trks)
        n_obs = sum(np.asarray(t).shape[0] for t in trks)
        rows0 = 2 * n_obs - 3 * n_trks
        cols = P.shape[1]
        self.jac = np.zeros((rows0, cols))
        self.cov_m_diag = np.ones(rows0)
        self.res = np.zeros(rows0)
        self.H1 = np.zeros((3 * n_trks, cols))
        self.H2 = np.zeros((3 * n_trks, 3 * n_trks))
        self.r1 = np.zeros(3 * n_trks)
        self.features = np.zeros(3 * n_trks)
        self.inlier = np.zeros(n_trks, dtype=bool)
        self.gamma = np.full(n_trks, np.nan)
        self.chi = np.full(n_trks, np.nan)
        var_img = sigma_img * sigma_img
        tri = Triangulation(quats, poss, max_iter, term) if n_trks else None  # shared, vio_updater.cpp:280
        row_h = 0
        n_p = len(quats)
        for j, trk in enumerate(trks):
            trk = np.asarray(trk, dtype=float)
            L = trk.shape[0]
            h_j = np.zeros((2 * L, cols))
            Hf_j = np.zeros((2 * L, 3))
            res_j = np.zeros(2 * L)
            a, b, r = tri.triangulate_gn(trk)
            R_n = rot(quats[-1])
            p_n = poss[-1]
            G_p_fj = 1.0 / r * R_n @ np.array([a, b, 1.0]) + p_n
            for i in range(L):
                pos = n_p - L + i
                R = rot(quats[pos])
                cp = R.T @ (G_p_fj - poss[pos])
                res_j[2 * i] = trk[i, 0] - cp[0] / cp[2]
                res_j[2 * i + 1] = trk[i, 1] - cp[1] / cp[2]
                if i == L - 1:  # :133-143
                    Hf_j[2 * i, 0] = 1.0
                    Hf_j[2 * i + 1, 1] = 1.0
                else:
                    J_i = _vis_jac(cp)
                    J_att = J_i @ skew(cp)
                    J_pos = -J_i @ R.T
                    J_anchor_att = -1.0 / r * J_i @ R.T @ R_n @ skew(np.array([a, b, 1.0]))
                    J_anchor_pos = -J_pos
                    Hf_j[2 * i:2 * i + 2] = 1.0 / r * J_i @ R.T @ R_n @ _mat_ivd(a, b, r)
                    c = K_CORE + pos * 3
                    h_j[2 * i:2 * i + 2, c:c + 3] = J_pos
                    c += n_poses_max * 3
                    h_j[2 * i:2 * i + 2, c:c + 3] = J_att
                    c = K_CORE + (n_p - 1) * 3
                    h_j[2 * i:2 * i + 2, c:c + 3] = J_anchor_pos
                    c += n_poses_max * 3
                    h_j[2 * i:2 * i + 2, c:c + 3] = J_anchor_att
            q = householder_q(Hf_j)
            A = q[:, 3:]
            U = q[:, :3]
            res0 = A.T @ res_j
            h0 = A.T @ h_j
            self.H1[3 * j:3 * j + 3] = U.T @ h_j
            self.H2[3 * j:3 * j + 3, 3 * j:3 * j + 3] = U.T @ Hf_j
            self.r1[3 * j:3 * j + 3] = U.T @ res_j
            self.features[3 * j:3 * j + 3] = (a, b, r)
            S = h0 @ P @ h0.T + var_img * np.eye(2 * L - 3)
            gamma = float(res0 @ np.linalg.inv(S) @ res0)
            chi = chi2_quantile(0.95, 2 * L - 3)
            self.gamma[j], self.chi[j] = gamma, chi
            if gamma < chi:
                self.inlier[j] = True
                self.jac[row_h:row_h + 2 * L - 3] = h0
                self.res[row_h:row_h + 2 * L - 3] = res0
                self.cov_m_diag[row_h:row_h + 2 * L - 3] = var_img
                row_h += 2 * L - 3
        self.rows_used = row_h
