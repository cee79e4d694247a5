"""Covariance-intersection fusion (oracle; test infrastructure only).

reference: src/x/ekf/ci.cpp, src/x/vio/multi_slam_update.cpp, src/x/ekf/simple_state.cpp,
           src/x/vio/msckf_update.cpp:88-139,175-279 (MULTI_UAV build).
Only the fixed-weight branch (0 < w <= 1) is restated; the w < 0 branch calls NLopt 2.7.1 LN_COBYLA
(third-party, CMakeLists.txt:135, ci.cpp:143-190) on singular inverses and stays out of scope.
"""
from dataclasses import dataclass, field

import numpy as np

from .quat import rot, skew
from .state import K_CORE
from .state_manager import _mat_ivd
from .triangulation import Triangulation
from .updates import chi2_quantile, global_feature_position, householder_q, msckf_track_jacobians


def _check_w(w_other):
    if w_other > 1.0 or w_other == 0 or w_other < -1:  # ci.cpp:59-62, 98-101
        raise RuntimeError("The CI weights must be lower than 1.0 and larger than 0.0")
    if w_other < 0.0:
        raise NotImplementedError("NLopt-optimised CI weights are out of scope (SURVEY.md 8f-4)")


def fuse_ci_pair(cov_a, H_a, cov_b, H_b, w_other):
    """reference: ci.cpp:94-127.  Returns (S, w_result)."""
    _check_w(w_other)
    P_a = H_a @ cov_a @ H_a.T
    P_b = H_b @ cov_b @ H_b.T
    w_result = 1.0 / (1.0 - w_other)
    S = (1.0 / (1.0 - w_other)) * P_a + (1.0 / w_other) * P_b
    return S, w_result


def fuse_ci_multi(cov_a, H_curr, covs, Hs, w_other):
    """reference: ci.cpp:49-92 (k-agent form).  Returns (S, w_result)."""
    _check_w(w_other)
    w0 = 1.0 - len(Hs) * w_other
    S = (1.0 / w0) * H_curr @ cov_a @ H_curr.T
    for H, c in zip(Hs, covs):
        S = S + (1.0 / w_other) * H @ c @ H.T
    return S, 1.0 / w0


@dataclass
class SimpleState:
    """Peer snapshot = the inter-agent wire payload.  reference: include/x/ekf/simple_state.h:30-75."""
    dynamic_state: np.ndarray
    positions_state: np.ndarray
    orientations_state: np.ndarray
    features_state: np.ndarray
    cov: np.ndarray
    anchor_idxs: list
    translation: np.ndarray = field(default_factory=lambda: np.zeros(3))

    def n_poses_max(self):
        return self.positions_state.size // 3

    def camera_attitudes(self):  # simple_state.cpp:34-49
        return [self.orientations_state[4 * i:4 * i + 4] for i in range(self.n_poses_max())]

    def camera_positions(self):  # simple_state.cpp:51-65
        return [self.positions_state[3 * i:3 * i + 3] + self.translation for i in range(self.n_poses_max())]


@dataclass
class SlamMatch:
    """reference: include/x/vision/types.h:102-116."""
    state: SimpleState
    current_feature_id: int
    received_feature_id: int


class MultiSlamUpdate:
    """reference: multi_slam_update.cpp:21-246.  Produces the (S, P_j, H, res) lists consumed by applyCI."""

    def __init__(self, quats, poss, feature_states, anchor_idxs, P, n_poses_max, sigma_landmark, matches, ci_slam_w):
        self.S_list, self.P_list, self.H_list, self.res_list = [], [], [], []
        self.gamma, self.inlier = [], []
        var_lm = sigma_landmark * sigma_landmark
        for m in matches:
            o = m.state
            self._one(quats, poss, feature_states, anchor_idxs[m.current_feature_id], m.current_feature_id, P,
                      n_poses_max, o.camera_attitudes(), o.camera_positions(), o.features_state,
                      o.anchor_idxs[m.received_feature_id], m.received_feature_id, o.cov, o.n_poses_max(),
                      var_lm, ci_slam_w)

    @staticmethod
    def jacobian(quats, poss, feats, anchor, fid, n_poses_max, cols, sign):
        """3 x N world-point Jacobian of one agent (multi_slam_update.cpp:116-203)."""
        a, b, r = feats[3 * fid:3 * fid + 3]
        if anchor < 0:
            raise RuntimeError("anchor_idx < 0")
        if r == 0:
            raise RuntimeError("rho = 0")
        R_a = rot(quats[anchor])
        G_p_f = (1.0 / r) * R_a @ np.array([a, b, 1.0]) + poss[anchor]
        h = np.zeros((3, cols))
        c = K_CORE + anchor * 3
        h[:, c:c + 3] = sign * np.eye(3)
        c += n_poses_max * 3
        h[:, c:c + 3] = sign * (-(1.0 / r) * R_a @ skew(np.array([a, b, 1.0])))
        c = K_CORE + (n_poses_max * 2 + fid) * 3
        h[:, c:c + 3] = sign * ((1.0 / r) * R_a @ _mat_ivd(a, b, r))
        return h, G_p_f

    def _one(self, quats, poss, feats, anchor, fid, P, M, o_quats, o_poss, o_feats, o_anchor, o_fid, o_P, o_M,
             var_lm, ci_slam_w):
        h_j, G_p_f = self.jacobian(quats, poss, feats, anchor, fid, M, P.shape[1], +1.0)
        oh_j, oG_p_f = self.jacobian(o_quats, o_poss, o_feats, o_anchor, o_fid, o_M, o_P.shape[1], -1.0)
        res_j = -G_p_f + oG_p_f
        r_j = var_lm * np.eye(3)
        S_gate = h_j @ P @ h_j.T + oh_j @ o_P @ oh_j.T + r_j
        gamma = float(res_j @ np.linalg.inv(S_gate) @ res_j)
        chi = chi2_quantile(0.9, 3)
        self.gamma.append(gamma)
        self.inlier.append(gamma < chi)
        if gamma < chi:
            S_j, w_result = fuse_ci_pair(P, h_j, o_P, oh_j, ci_slam_w)
            S_j = S_j + r_j
            P_j = P.copy()
            for c in (K_CORE + anchor * 3, K_CORE + anchor * 3 + M * 3, K_CORE + (M * 2 + fid) * 3):
                P_j[c:c + 3, c:c + 3] *= w_result  # only the three diagonal 3x3 blocks, :229-239
            self.H_list.append(h_j)
            self.S_list.append(S_j)
            self.res_list.append(res_j)
            self.P_list.append(P_j)


@dataclass
class MsckfMatch:
    """reference: include/x/vision/types.h:83-100."""
    state: SimpleState
    id_current_track: int
    received_track: np.ndarray  # (L_peer, 2)


def consume_matches(matches, track_id):
    """The match-collection loop of preProcessOneTrack (msckf_update.cpp:96-139) on the SHARED, mutable list:
    matches of `track_id` are moved out of `matches`.  The loop bound is re-evaluated while the list shrinks and the
    index is corrected by the number of erasures, so of several matches sitting at the tail of the list only some are
    visited (e.g. [a(t), b(t)] consumes a only) -- restated as written."""
    mine = []
    corrected = 0
    i = 0
    while i < len(matches):
        if matches[i - corrected].id_current_track == track_id:
            mine.append(matches.pop(i - corrected))
            corrected += 1
        i += 1
    return mine


def multi_msckf_one_track(track, mine, quats, poss, P, n_poses_max, sigma_img, ci_msckf_w, max_iter=10, term=1e-5):
    """Joint multi-agent MSCKF block for ONE own track (msckf_update.cpp:65-281, MULTI_UAV build).

    `mine` = the MsckfMatch entries consume_matches() moved out of the shared list for this track.
    Returns dict(own=(inlier, gamma), jac0, res0, multi=None | (S_j, P_j, h_j, res_pf))."""
    var_img = sigma_img * sigma_img
    track = np.asarray(track, dtype=float)
    L = track.shape[0]
    tmp_q, tmp_p, tmp_trk = [], [], []
    sizes = [P.shape[1]]
    for m in mine:  # :96-139 -- peers first, own poses last
        Lp = m.received_track.shape[0]
        tmp_p += m.state.camera_positions()[-Lp:]
        tmp_q += m.state.camera_attitudes()[-Lp:]
        tmp_trk.append(m.received_track)
        sizes.append(m.state.cov.shape[0])
    tmp_p += poss[len(poss) - L:]
    tmp_q += quats[len(quats) - L:]
    tmp_trk.append(track)
    tmp_trk = np.vstack(tmp_trk)
    tri = Triangulation(tmp_q, tmp_p, max_iter, term)
    ivd = tri.triangulate_gn(tmp_trk)
    G_p_fj = global_feature_position(ivd, tmp_q[-1], tmp_p[-1])

    def project(trk, q_l, p_l, M, cols):
        out = msckf_track_jacobians(trk, q_l, p_l, M, cols, G_p_fj)
        if out is None:
            return None
        jac_j, Hf_j, res_j = out
        q = householder_q(Hf_j)
        return jac_j, Hf_j, res_j, q[:, :3], q[:, 3:]

    own = project(track, quats, poss, n_poses_max, P.shape[1])
    if own is None:
        return dict(own=(False, np.nan), multi=None, ivd=ivd, G_p_f=G_p_fj, n_matched=len(mine))
    jac_j, Hf_j, res_j, A_up, A = own
    res0, jac0 = A.T @ res_j, A.T @ jac_j
    S = jac0 @ P @ jac0.T + var_img * np.eye(2 * L - 3)
    gamma = float(res0 @ np.linalg.inv(S) @ res0)
    inl = gamma < chi2_quantile(0.95, 2.0 * L - 3.0)
    result = dict(own=(inl, gamma), multi=None, jac0=jac0, res0=res0, ivd=ivd, G_p_f=G_p_fj, n_matched=len(mine))
    if not (inl and mine):
        return result
    k = len(mine)
    tot_cols = sum(sizes)
    jac_x_pf = np.zeros((3 * (k + 1), tot_cols))
    jac_pf = np.zeros((3 * (k + 1), 3))
    res_pf = np.zeros(3 * (k + 1))
    jac_x_pf[0:3, 0:sizes[0]] = A_up.T @ jac_j
    jac_pf[0:3] = A_up.T @ Hf_j
    res_pf[0:3] = A_up.T @ res_j
    col = sizes[0]
    for i, m in enumerate(mine):
        o = m.state
        pr = project(m.received_track, o.camera_attitudes(), o.camera_positions(), o.n_poses_max(), o.cov.shape[0])
        if pr is not None:
            oj, oHf, ores, oA_up, _ = pr
            jac_x_pf[3 * (i + 1):3 * (i + 2), col:col + sizes[i + 1]] = oA_up.T @ oj
            jac_pf[3 * (i + 1):3 * (i + 2)] = oA_up.T @ oHf
            res_pf[3 * (i + 1):3 * (i + 2)] = oA_up.T @ ores
        col += sizes[i + 1]
    q = householder_q(jac_pf)  # nullSpaceProjection, :494-501
    A2 = q[:, 3:]
    jac_x_pf = A2.T @ jac_x_pf
    res_pf = A2.T @ res_pf
    h_j = jac_x_pf[:, :sizes[0]]
    S_j = h_j @ P @ h_j.T
    Hs = []
    col = sizes[0]
    for i, m in enumerate(mine):
        Hs.append(jac_x_pf[:, col:col + sizes[i + 1]])
        col += sizes[i + 1]
        S_j = S_j + Hs[i] @ m.state.cov @ Hs[i].T
    S_j = S_j + var_img * np.eye(S_j.shape[0])
    gamma_m = float(res_pf @ np.linalg.inv(S_j) @ res_pf)
    chi_m = chi2_quantile(0.95, 2.0 * tmp_trk.shape[0] - 3.0)
    result["multi_gate"] = (gamma_m, chi_m)
    if gamma_m < chi_m:
        S_ci, w_result = fuse_ci_multi(P, h_j, [m.state.cov for m in mine], Hs, ci_msckf_w)
        S_ci = S_ci + var_img * np.eye(S_ci.shape[0])
        P_j = P.copy()
        n_p = len(quats)
        for i in range(L):  # :258-267
            c = K_CORE + (n_p - L + i) * 3
            P_j[c:c + 3, c:c + 3] *= w_result
            c += n_poses_max * 3
            P_j[c:c + 3, c:c + 3] *= w_result
        result["multi"] = (S_ci, P_j, h_j, res_pf)
    return result


class MultiMsckfUpdate:
    """MsckfUpdate of the MULTI_UAV build (msckf_update.cpp:27-63 with `tracks_matches`): same stacked (jac, res,
    cov_m_diag) as the single-agent class -- except that matched tracks are triangulated jointly with the peers'
    observations -- plus the (S, P_j, H, res) lists consumed by Updater::applyCI.  `track_ids[j]` is what
    MsckfMatch.id_current_track is compared with (Track::getId, :98); `matches` is mutated like the reference's."""

    def __init__(self, trks, track_ids, quats, poss, P, n_poses_max, sigma_img, matches, ci_msckf_w):
        n_trks = len(trks)
        n_obs = sum(np.asarray(t).shape[0] for t in trks)
        rows = 2 * n_obs - 3 * n_trks
        cols = P.shape[1]
        self.jac = np.zeros((rows, cols))
        self.cov_m_diag = np.ones(rows)
        self.res = np.zeros(rows)
        self.inlier = np.zeros(n_trks, dtype=bool)
        self.gamma = np.full(n_trks, np.nan)
        self.G_p_f = np.full((n_trks, 3), np.nan)
        self.n_matched = np.zeros(n_trks, dtype=int)
        self.multi_gate = {}
        self.S_list, self.P_list, self.H_list, self.res_list = [], [], [], []
        var_img = sigma_img * sigma_img
        row_h = 0
        for j, trk in enumerate(trks):
            trk = np.asarray(trk, dtype=float)
            L = trk.shape[0]
            mine = consume_matches(matches, track_ids[j])
            out = multi_msckf_one_track(trk, mine, quats, poss, P, n_poses_max, sigma_img, ci_msckf_w)
            self.G_p_f[j] = out["G_p_f"]
            self.n_matched[j] = out["n_matched"]
            inl, gamma = out["own"]
            self.inlier[j], self.gamma[j] = inl, gamma
            if inl:
                self.jac[row_h:row_h + 2 * L - 3] = out["jac0"]
                self.res[row_h:row_h + 2 * L - 3] = out["res0"]
                self.cov_m_diag[row_h:row_h + 2 * L - 3] = var_img
                row_h += 2 * L - 3
            if "multi_gate" in out:
                self.multi_gate[j] = out["multi_gate"]
            if out["multi"] is not None:
                S_j, P_j, h_j, res_pf = out["multi"]
                self.S_list.append(S_j); self.P_list.append(P_j); self.H_list.append(h_j); self.res_list.append(res_pf)
        self.rows_used = row_h


# ---- compressed payload (test infrastructure for the multi-agent exchange) ------------------------
def pack_payload(state, sm, n_features_max):
    """[8 header | 13 per feature: valid, G_p_f(3), h P h^T(9)] -- what xb_ci_pack computes on the device."""
    out = np.zeros(8 + 13 * max(1, n_features_max))
    out[0], out[1] = sm.n_features, state.time
    quats, poss = sm.camera_attitudes(state), sm.camera_positions(state)
    for f in range(sm.n_features):
        h, G = MultiSlamUpdate.jacobian(quats, poss, state.f_array, sm.anchor_idxs[f], f, sm.n_poses_max,
                                        state.cov.shape[1], +1.0)
        o = out[8 + 13 * f:8 + 13 * f + 13]
        o[0] = 1.0
        o[1:4] = G
        o[4:13] = (h @ state.cov @ h.T).ravel()
    return out


def multi_slam_from_payload(quats, poss, feature_states, anchor_idxs, P, n_poses_max, sigma_landmark, gathered, matches,
                            ci_slam_w):
    """MultiSlamUpdate restated on gathered payload slots; `matches` = (peer_slot, current_fid, received_fid)."""
    _check_w(ci_slam_w)
    var_lm = sigma_landmark * sigma_landmark
    out = dict(S=[], P=[], H=[], res=[], gamma=[], inlier=[])
    for peer, cur, rcv in matches:
        o = gathered[peer][8 + 13 * rcv:8 + 13 * rcv + 13]
        h_j, G_p_f = MultiSlamUpdate.jacobian(quats, poss, feature_states, anchor_idxs[cur], cur, n_poses_max, P.shape[1], +1.0)
        Mo, Mp = h_j @ P @ h_j.T, o[4:13].reshape(3, 3)
        res_j = -G_p_f + o[1:4]
        gamma = float(res_j @ np.linalg.inv(Mo + Mp + var_lm * np.eye(3)) @ res_j)
        inl = gamma < chi2_quantile(0.9, 3)
        out["gamma"].append(gamma)
        out["inlier"].append(inl)
        if inl:
            w_res = 1.0 / (1.0 - ci_slam_w)
            S_j = w_res * Mo + (1.0 / ci_slam_w) * Mp + var_lm * np.eye(3)
            P_j = P.copy()
            a = anchor_idxs[cur]
            for c in (K_CORE + a * 3, K_CORE + a * 3 + n_poses_max * 3, K_CORE + (n_poses_max * 2 + cur) * 3):
                P_j[c:c + 3, c:c + 3] *= w_res
            out["S"].append(S_j); out["P"].append(P_j); out["H"].append(h_j); out["res"].append(res_j)
    return out


def pack_pose_payload(state, n_poses_max):
    """[8 header | 3M camera positions | 4M attitudes | 6M x 6M pose block of the covariance] -- what xb_ci_pack_poses
    writes on the device: everything a peer contributes to the MSCKF-MSCKF block (msckf_update.cpp:175-279)."""
    M = n_poses_max
    n6 = 6 * M
    out = np.zeros(8 + 7 * M + n6 * n6)
    out[0], out[1], out[2] = 1.0, state.time, M
    out[8:8 + 3 * M] = state.p_array
    out[8 + 3 * M:8 + 7 * M] = state.q_array
    out[8 + 7 * M:] = state.cov[K_CORE:K_CORE + n6, K_CORE:K_CORE + n6].ravel()
    return out


def peer_from_pose_payload(payload, n_features_max=0):
    """SimpleState whose covariance holds only the pose block (all that MultiMsckfUpdate reads of a peer)."""
    M = int(payload[2])
    n6 = 6 * M
    N = K_CORE + n6 + 3 * n_features_max
    cov = np.zeros((N, N))
    cov[K_CORE:K_CORE + n6, K_CORE:K_CORE + n6] = payload[8 + 7 * M:].reshape(n6, n6)
    return SimpleState(np.zeros(16), payload[8:8 + 3 * M].copy(), payload[8 + 3 * M:8 + 7 * M].copy(),
                       np.zeros(3 * n_features_max), cov, [-1] * n_features_max)
