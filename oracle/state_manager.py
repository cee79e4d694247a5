"""Sliding-window bookkeeping on state and covariance (oracle; test infrastructure only).

reference: src/x/vio/state_manager.cpp (dense J P J^T formulation kept on purpose: this is what the reference executes).
"""
import numpy as np

from .quat import rot, skew, qconj
from .state import K_CORE


def _mat_ivd(alpha, beta, rho):
    """The 3x3 `mat` used by the inverse-depth Jacobians (slam_update.cpp:153-157, state_manager.cpp:433-436)."""
    m = np.eye(3)
    m[0, 2] = -alpha / rho
    m[1, 2] = -beta / rho
    m[2, 2] = -1.0 / rho
    return m


class StateManager:
    def __init__(self, n_poses_max, n_features_max):
        self.n_poses_max = n_poses_max
        self.n_features_max = n_features_max
        self.clear()

    def clear(self):  # state_manager.cpp:22-29
        self.n_poses = 0
        self.n_features = 0
        self.anchor_idxs = [-1] * self.n_features_max
        self.filled_before = False

    def copy(self):
        o = StateManager(self.n_poses_max, self.n_features_max)
        o.n_poses, o.n_features = self.n_poses, self.n_features
        o.anchor_idxs, o.filled_before = list(self.anchor_idxs), self.filled_before
        return o

    # ------------------------------------------------------------------ manage
    def manage(self, state, del_feat_idx):
        """reference: state_manager.cpp:31-149."""
        M, F = self.n_poses_max, self.n_features_max
        att = state.q_array.copy()
        pos = state.p_array.copy()
        cae = state.camera_orientation()
        cpe = state.camera_position()
        cov = state.cov.copy()
        feats = state.f_array.copy()

        for idx in sorted(del_feat_idx, reverse=True):  # :52-112 (largest index first)
            n1 = self.n_features - idx - 1
            feats[idx * 3:(idx + n1) * 3] = feats[(idx + 1) * 3:(idx + 1 + n1) * 3].copy()
            feats[(self.n_features - 1) * 3:(self.n_features - 1) * 3 + 3] = 0.0
            del self.anchor_idxs[idx]
            self.anchor_idxs.append(-1)
            n = cov.shape[0]
            idx0 = K_CORE + M * 6 + idx * 3
            idx1 = idx0 + 3
            dim0 = (F - idx - 1) * 3
            cols_after = cov[:, idx1:idx1 + dim0].copy()
            cols_after[idx0:idx0 + dim0, :] = cols_after[idx1:idx1 + dim0, :].copy()
            rows_after = cov[idx1:idx1 + dim0, :idx0].copy()
            cov[:, idx0:idx0 + dim0] = cols_after
            cov[idx0:idx0 + dim0, :idx0] = rows_after
            cov[:, n - 3:] = 0.0
            cov[n - 3:, :] = 0.0
            self.n_features -= 1

        if self.n_poses == M:  # :119-125
            cov = self.reparametrize_features(att, pos, feats, cov)
            cov = self.slide_window(att, pos, cov)

        att[self.n_poses * 4:self.n_poses * 4 + 4] = cae  # :133
        pos[self.n_poses * 3:self.n_poses * 3 + 3] = cpe
        cov = self.augment_covariance(state, self.n_poses, cov)
        self.n_poses += 1
        state.cov, state.q_array, state.p_array, state.f_array = cov, att, pos, feats

    # ------------------------------------------------------------------ augment
    def augment_covariance(self, state, pos, cov):
        """reference: state_manager.cpp:273-349."""
        M, F = self.n_poses_max, self.n_features_max
        n = K_CORE + 6 * M + 3 * F
        J = np.eye(n) if self.filled_before else np.zeros((n, n))
        k = K_CORE + (pos + 1) * 3
        J[:k, :k] = np.eye(k)
        a0 = K_CORE + 3 * M
        J[a0:a0 + (pos + 1) * 3, a0:a0 + (pos + 1) * 3] = np.eye((pos + 1) * 3)
        f0 = K_CORE + 6 * M
        J[f0:f0 + 3 * self.n_features, f0:f0 + 3 * self.n_features] = np.eye(3 * self.n_features)
        rp = K_CORE + pos * 3
        ra = a0 + pos * 3
        J[rp:rp + 3, 0:3] = np.eye(3)
        J[rp:rp + 3, 6:9] = -rot(state.q) @ skew(state.p_ic)
        J[ra:ra + 3, 6:9] = rot(qconj(state.q_ic))
        P = cov.copy()
        P[rp:rp + 3, :] = 0.0
        P[ra:ra + 3, :] = 0.0
        P[:, rp:rp + 3] = 0.0
        P[:, ra:ra + 3] = 0.0
        if pos + 1 == M:
            self.filled_before = True
        return J @ P @ J.T

    # ------------------------------------------------------------------ reparametrize
    def reparametrize_features(self, atts_old, poss_old, features, cov):
        """reference: state_manager.cpp:351-482 (eq. 38, Li RSS-2012 supplementals). `features` updated in place."""
        M, F = self.n_poses_max, self.n_features_max
        q_old = atts_old[0:4]
        p_old = poss_old[0:3]
        idx_to_chg = [i for i in range(self.n_features) if self.anchor_idxs[i] == 0]
        n = K_CORE + 6 * M + 3 * F
        J = np.eye(n)
        idx1 = M - 1
        q_new = atts_old[4 * idx1:4 * idx1 + 4]
        p_new = poss_old[3 * idx1:3 * idx1 + 3]
        R_new_T = rot(q_new).T
        R_old = rot(q_old)
        for j in idx_to_chg:
            a_o, b_o, r_o = features[3 * j:3 * j + 3]
            new_params = R_new_T @ (-p_new + p_old + 1.0 / r_o * R_old @ np.array([a_o, b_o, 1.0]))
            r_n = 1.0 / new_params[2]
            a_n = new_params[0] * r_n
            b_n = new_params[1] * r_n
            features[3 * j:3 * j + 3] = (a_n, b_n, r_n)
            self.anchor_idxs[j] = idx1
            J_att_old = -1.0 / r_o * R_new_T @ R_old @ skew(np.array([a_o, b_o, 1.0]))
            J_att_new = skew(new_params)
            J_pos_old = R_new_T
            J_pos_new = -R_new_T
            J_feat_old = 1.0 / r_o * R_new_T @ R_old @ _mat_ivd(a_o, b_o, r_o)
            A = np.zeros((3, n))
            c = K_CORE + idx1 * 3
            A[:, c:c + 3] = J_pos_new
            c += 3 * M
            A[:, c:c + 3] = J_att_new
            c = K_CORE
            A[:, c:c + 3] = J_pos_old
            c += 3 * M
            A[:, c:c + 3] = J_att_old
            c = K_CORE + 6 * M + 3 * j
            A[:, c:c + 3] = J_feat_old
            mat = np.eye(3)
            mat[0, 2], mat[1, 2], mat[2, 2] = -a_n, -b_n, -r_n
            J[c:c + 3, :] = r_n * mat @ A
        return J @ cov @ J.T

    # ------------------------------------------------------------------ slide
    def slide_window(self, atts, poss, cov):
        """reference: state_manager.cpp:484-537 (pure 0/1 permutation, dense in the reference)."""
        M, F = self.n_poses_max, self.n_features_max
        atts[:(M - 1) * 4] = atts[4:].copy()
        poss[:(M - 1) * 3] = poss[3:].copy()
        atts[(M - 1) * 4:] = 0.0
        poss[(M - 1) * 3:] = 0.0
        n = cov.shape[0]
        L = np.zeros((n, n))
        L[:K_CORE, :K_CORE] = np.eye(K_CORE)
        if F:
            f0 = K_CORE + 6 * M
            L[f0:, f0:] = np.eye(3 * F)
        Rm = L.copy()
        m3 = (M - 1) * 3
        L[K_CORE:K_CORE + m3, K_CORE + 3:K_CORE + 3 + m3] = np.eye(m3)
        L[K_CORE + 3 * M:K_CORE + 3 * M + m3, K_CORE + 3 + 3 * M:K_CORE + 3 + 3 * M + m3] = np.eye(m3)
        Rm[K_CORE + 3:K_CORE + 3 + m3, K_CORE:K_CORE + m3] = np.eye(m3)
        Rm[K_CORE + 3 + 3 * M:K_CORE + 3 + 3 * M + m3, K_CORE + 3 * M:K_CORE + 3 * M + m3] = np.eye(m3)
        cov = L @ cov @ Rm
        for i in range(self.n_features):
            self.anchor_idxs[i] -= 1
        self.n_poses -= 1
        return cov

    # ------------------------------------------------------------------ feature init
    def init_msckf_slam_features(self, state, H1, H2, r1, features, correction, sigma_img):
        """reference: state_manager.cpp:151-174 (Li 2012)."""
        P = state.cov
        H2_inv = np.linalg.inv(H2)
        H2_inv_H1 = H2_inv @ H1
        new_f = features - H2_inv_H1 @ correction + H2_inv @ r1
        var = sigma_img * sigma_img
        P_cross = -H2_inv_H1 @ P
        P_diag = H2_inv_H1 @ P @ H2_inv_H1.T + var * H2_inv @ H2_inv.T
        self.add_feature_states(state, new_f, P_diag, P_cross)

    def init_standard_slam_features(self, state, new_f, sigma_img, sigma_rho_0):
        """reference: state_manager.cpp:176-198."""
        n_new = new_f.size
        n = state.cov.shape[0]
        P_cross = np.zeros((n_new, n))
        P_diag = sigma_img * sigma_img * np.eye(n_new)
        for i in range(n_new // 3):
            P_diag[3 * i + 2, 3 * i + 2] = sigma_rho_0 * sigma_rho_0
        self.add_feature_states(state, new_f, P_diag, P_cross)

    def add_feature_states(self, state, new_f, cov, cross):
        """reference: state_manager.cpp:200-227."""
        n_new = new_f.size
        feats = state.f_array.copy()
        assert self.n_features < self.n_features_max
        feats[self.n_features * 3:self.n_features * 3 + n_new] = new_f
        state.f_array = feats
        P = state.cov
        ns = K_CORE + 6 * self.n_poses_max + 3 * self.n_features
        P[ns:ns + n_new, :] = cross
        P[:, ns:ns + n_new] = cross.T
        P[ns:ns + n_new, ns:ns + n_new] = cov
        for i in range(n_new // 3):
            self.anchor_idxs[self.n_features + i] = self.n_poses - 1
        self.n_features += n_new // 3

    # ------------------------------------------------------------------ lists
    def camera_attitudes(self, state, max_size=0):  # state_manager.cpp:539-566
        size_out = min(max_size, self.n_poses) if max_size > 0 else self.n_poses
        start = self.n_poses - size_out
        return [state.q_array[4 * i:4 * i + 4].copy() for i in range(start, self.n_poses)]

    def camera_positions(self, state):  # state_manager.cpp:568-584
        return [state.p_array[3 * i:3 * i + 3].copy() for i in range(self.n_poses)]
