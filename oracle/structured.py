"""TEST / BENCH INFRASTRUCTURE (oracle): the STRUCTURED formulation of one visual update on the CPU, vectorised numpy.

The device does not execute the reference's dense algebra (136.6 GFLOP per cfg-2 update) but the structured
form of DESIGN.md section 2 (projector gate on 2L rows, Gram compression, sparse SLAM rows, Cholesky instead of the explicit
inverse: ~1-2 GFLOP).  This module runs that same formulation on the host (numpy + the BLAS/LAPACK numpy is linked with)
so that `bench.py` can report how much of the GPU-vs-reference ratio is the algorithm and how much is the hardware:
`cpu_baseline.structured`.  It is NOT the product and NOT the parity oracle (that is the dense restatement in this
package, pinned to the compiled reference); bench.py checks its state correction against the device's.

Scope: the steady-state update of the bench (every MSCKF track spans the whole window, no new SLAM features, plain
MSCKF Jacobians = xb_config.oc_projection 0), stages after StateManager::manage: triangulation (two-view DLT +
Gauss-Newton, triangulation.cpp:102-206), Jacobians (msckf_update.cpp:328-417), projector gate, Gram compression
(vio_updater.cpp:487-512), SLAM rows + gates (slam_update.cpp:49-214), Kalman update (updater.cpp:117-141).
"""
import time

import numpy as np

K_CORE = 15


def _rot(q):
    """q (n, 4) as (x, y, z, w) -> normalised rotation matrices (n, 3, 3)."""
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _skew(v):
    """(..., 3) -> (..., 3, 3)"""
    S = np.zeros(v.shape + (3,))
    S[..., 0, 1], S[..., 0, 2] = -v[..., 2], v[..., 1]
    S[..., 1, 0], S[..., 1, 2] = v[..., 2], -v[..., 0]
    S[..., 2, 0], S[..., 2, 1] = -v[..., 1], v[..., 0]
    return S


def _vis_jac(cp):
    """(..., 3) -> (..., 2, 3)"""
    J = np.zeros(cp.shape[:-1] + (2, 3))
    iz = 1.0 / cp[..., 2]
    J[..., 0, 0] = iz; J[..., 0, 2] = -cp[..., 0] * iz * iz
    J[..., 1, 1] = iz; J[..., 1, 2] = -cp[..., 1] * iz * iz
    return J


def structured_update(p_array, q_array, f_array, anchors, P, Z, slam_obs, slam_len, n_poses, M, F, sigma_img, chi2_95, chi2_90,
                      force_inliers=None):
    """One update from the post-manage work state.  Z: (K, L, 2) MSCKF tracks (L == n_poses), slam_obs: (F, 2) last
    observation of every SLAM feature, slam_len: (F,) track sizes.  force_inliers = (msckf mask, slam mask) overrides the
    gates (bench.py uses the device's masks when it compares the state correction).  Returns (delta, info) with per-stage
    seconds."""
    tm = {}
    t0 = time.perf_counter()
    N = P.shape[0]
    var = sigma_img * sigma_img
    ps = p_array.reshape(M, 3)[:n_poses]
    R = _rot(q_array.reshape(M, 4)[:n_poses])
    K, L = Z.shape[:2]
    i1 = n_poses - L
    Rw, pw = R[i1:], ps[i1:]
    Rl, pl = R[-1], ps[-1]
    # ---- triangulation
    Rt = Rw.transpose(0, 2, 1)
    Pm = np.concatenate([Rt, -np.einsum("iab,ib->ia", Rt, pw)[..., None]], axis=2)
    P1, P2 = Pm[0], Pm[-1]
    A = np.stack([Z[:, 0, 0:1] * P1[2] - P1[0], Z[:, 0, 1:2] * P1[2] - P1[1],
                  Z[:, -1, 0:1] * P2[2] - P2[0], Z[:, -1, 1:2] * P2[2] - P2[1]], axis=1)
    vt = np.linalg.svd(A)[2][:, 3, :]
    X3 = vt[:, :3] / vt[:, 3:]
    c2 = X3 @ P2[:, :3].T + P2[:, 3]
    th = np.stack([c2[:, 0] / c2[:, 2], c2[:, 1] / c2[:, 2], 1.0 / c2[:, 2]], axis=1)   # alpha, beta, rho
    dR = np.einsum("iba,bc->iac", Rw, Rl)                  # R_i^T R_l
    dp = np.einsum("iba,ib->ia", Rw, pl - pw)              # R_i^T (p_l - p_i)
    j0 = np.stack([dR[:, :, 0], dR[:, :, 1], dp], axis=2)  # (L, 3 rows, 3 params)
    last, cur = np.full(K, 1000.0), np.full(K, 100.0)
    active = np.ones(K, dtype=bool)
    for it in range(1, 11):
        active &= (last - cur) > 1e-5
        if not active.any():
            break
        a = np.flatnonzero(active)
        h = dR[None, :, :, 0] * th[a, None, None, 0] + dR[None, :, :, 1] * th[a, None, None, 1] + dR[None, :, :, 2] \
            + th[a, None, None, 2] * dp[None]
        r = Z[a] - h[..., :2] / h[..., 2:]
        ih = 1.0 / h[..., 2]
        J = np.empty((len(a), L, 2, 3))
        J[:, :, 0] = -ih[..., None] * j0[None, :, 0] + (h[..., 0] * ih * ih)[..., None] * j0[None, :, 2]
        J[:, :, 1] = -ih[..., None] * j0[None, :, 1] + (h[..., 1] * ih * ih)[..., None] * j0[None, :, 2]
        JtJ = np.einsum("klai,klaj->kij", J, J)
        Jtr = np.einsum("klai,kla->ki", J, r)
        th[a] -= np.linalg.solve(JtJ, Jtr[..., None])[..., 0]
        last[a] = cur[a]
        cur[a] = np.sqrt(np.einsum("kla,kla->k", r, r))
    Gf = (np.stack([th[:, 0], th[:, 1], np.ones(K)], axis=1) @ Rl.T) / th[:, 2:3] + pl
    tm["triangulation"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # ---- Jacobians
    cp = np.einsum("iba,kib->kia", Rw, Gf[:, None, :] - pw[None])
    res = Z - cp[..., :2] / cp[..., 2:]
    Ji = _vis_jac(cp)
    Jpos = -np.einsum("klab,lcb->klac", Ji, Rw)
    Jatt = Ji @ _skew(cp)
    J6 = np.concatenate([Jpos, Jatt], axis=-1)             # (K, L, 2, 6)
    U = np.linalg.qr((-Jpos).reshape(K, 2 * L, 3))[0]
    tm["jacobians_basis"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # ---- gate: X = J P J^T over the pose blocks, S = Pi X Pi + var I
    cols6 = np.concatenate([K_CORE + 3 * (i1 + np.arange(L))[:, None] + np.arange(3),
                            K_CORE + 3 * M + 3 * (i1 + np.arange(L))[:, None] + np.arange(3)], axis=1)   # (L, 6)
    Ps = 0.5 * (P + P.T)
    Pp = Ps[cols6[:, None, :, None], cols6[None, :, None, :]]                                            # (L_i, L_j, 6, 6)
    # T[k, i, a, (j, c)] = J6[k, i, a, :] Pp[i, :, j, c]: one (2K x 6)(6 x 6L) product per pose i;
    # X[k, (i, a), (j, d)] = T[k, i, a, j, :] . J6[k, j, d, :]: batched (2L x 6)(6 x 2) products
    T = np.empty((K, L, 2, L, 6))
    for i in range(L):
        T[:, i] = (J6[:, i].reshape(2 * K, 6) @ Pp[i].transpose(1, 0, 2).reshape(6, 6 * L)).reshape(K, 2, L, 6)
    X = np.matmul(T.reshape(K, 2 * L, L, 6).transpose(0, 2, 1, 3), J6.transpose(0, 1, 3, 2))      # (K, L_j, 2L, 2)
    X = X.transpose(0, 2, 1, 3).reshape(K, 2 * L, 2 * L)
    r2 = res.reshape(K, 2 * L)
    Ur = np.einsum("kru,kr->ku", U, r2)
    pr = r2 - np.einsum("kru,ku->kr", U, Ur)
    Y = X @ U
    Zm = np.einsum("kru,krv->kuv", U, Y)
    W = Y - 0.5 * U @ Zm
    S = X - U @ W.transpose(0, 2, 1) - W @ U.transpose(0, 2, 1)
    S[:, np.arange(2 * L), np.arange(2 * L)] += var
    gamma = np.einsum("kr,kr->k", pr, np.linalg.solve(S, pr[..., None])[..., 0])
    inl = gamma < chi2_95[2 * L - 3]
    if force_inliers is not None:
        inl = np.asarray(force_inliers[0], dtype=bool)
    tm["gate"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # ---- Gram compression: G = [J|r]^T [J|r] - [B|b]^T [B|b] over the inliers, G = R^T R
    n = 6 * M
    Wd = n + 1
    ki = np.flatnonzero(inl)
    J7 = np.concatenate([J6[ki], res[ki][..., None]], axis=-1)                                           # (Ki, L, 2, 7)
    blk = np.einsum("klai,klaj->lij", J7, J7)
    G = np.zeros((Wd, Wd))
    c7 = np.concatenate([cols6 - K_CORE, np.full((L, 1), n)], axis=1)                                     # (L, 7)
    for l in range(L):
        G[np.ix_(c7[l], c7[l])] += blk[l]
    Bf = np.zeros((len(ki), 3, Wd))
    Bp = np.einsum("klau,klac->kulc", U[ki].reshape(len(ki), L, 2, 3), J6[ki])                           # (Ki, 3, L, 6)
    Bf[:, :, (cols6 - K_CORE).reshape(-1)] = Bp.reshape(len(ki), 3, 6 * L)
    Bf[:, :, n] = Ur[ki]
    B2 = Bf.reshape(-1, Wd)
    G -= B2.T @ B2
    Gn = G[:n, :n]
    # guarded factorisation of the semi-definite Gram matrix: eigen-decomposition, non-positive directions dropped
    ev, Q = np.linalg.eigh(Gn)
    keep = ev > 1e-14 * ev.max()
    Rg = (Q[:, keep] * np.sqrt(ev[keep])).T                # rows: compressed measurement on the pose columns
    zg = (Q[:, keep] / np.sqrt(ev[keep])).T @ G[:n, n]
    tm["gram_compression"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # ---- SLAM rows and their 2x2 gates
    fs = f_array.reshape(F, 3)
    anc = np.asarray(anchors[:F])
    Ra, pa = R[anc], ps[anc]
    ab1 = np.stack([fs[:, 0], fs[:, 1], np.ones(F)], axis=1)
    Gs = np.einsum("fab,fb->fa", Ra, ab1) / fs[:, 2:3] + pa
    cps = (Gs - pl) @ Rl                                   # R_n^T (G - p_n)
    rs = slam_obs - cps[:, :2] / cps[:, 2:]
    Jis = _vis_jac(cps)
    JposS = -Jis @ Rl.T
    JattS = Jis @ _skew(cps)
    RtRa = np.einsum("ba,fbc->fac", Rl, Ra)
    JR = Jis @ RtRa
    JancA = -(JR @ _skew(ab1)) / fs[:, 2, None, None]
    m3 = np.zeros((F, 3, 3))
    m3[:, 0, 0] = 1.0; m3[:, 1, 1] = 1.0
    m3[:, 0, 2] = -fs[:, 0] / fs[:, 2]; m3[:, 1, 2] = -fs[:, 1] / fs[:, 2]; m3[:, 2, 2] = -1.0 / fs[:, 2]
    HfS = (JR @ m3) / fs[:, 2, None, None]
    Hs = np.zeros((F, 2, N))
    pos = n_poses - 1
    ar = np.arange(F)
    for c in range(3):
        Hs[:, :, K_CORE + 3 * pos + c] += JposS[:, :, c]
        Hs[:, :, K_CORE + 3 * M + 3 * pos + c] += JattS[:, :, c]
        Hs[ar, :, K_CORE + 3 * anc + c] += -JposS[:, :, c]
        Hs[ar, :, K_CORE + 3 * M + 3 * anc + c] += JancA[:, :, c]
        Hs[ar, :, K_CORE + 6 * M + 3 * ar + c] += HfS[:, :, c]
    same = anc == pos                                      # slam_update.cpp:120-130
    if same.any():
        Hs[same] = 0.0
        Hs[same, 0, K_CORE + 6 * M + 3 * ar[same]] = 1.0
        Hs[same, 1, K_CORE + 6 * M + 3 * ar[same] + 1] = 1.0
    # the rows are <= 15 wide: gather the 15 x 15 block of P per feature instead of the dense product
    cols15 = np.concatenate([K_CORE + 3 * pos + np.arange(3)[None].repeat(F, 0), K_CORE + 3 * M + 3 * pos + np.arange(3)[None].repeat(F, 0),
                             K_CORE + 3 * anc[:, None] + np.arange(3), K_CORE + 3 * M + 3 * anc[:, None] + np.arange(3),
                             K_CORE + 6 * M + 3 * ar[:, None] + np.arange(3)], axis=1)             # (F, 15)
    h15 = np.take_along_axis(Hs, cols15[:, None, :].repeat(2, 1), axis=2)                               # (F, 2, 15)
    if same.any():
        h15[same, :, :12] = 0.0      # duplicate columns (anchor == newest pose) would be counted twice
    P15 = P[cols15[:, :, None], cols15[:, None, :]]
    Ss = h15 @ P15 @ h15.transpose(0, 2, 1)
    Ss[:, 0, 0] += var; Ss[:, 1, 1] += var
    gs = np.einsum("fa,fa->f", rs, np.linalg.solve(Ss, rs[..., None])[..., 0])
    inl_s = gs < chi2_90[np.minimum(2 * np.asarray(slam_len), len(chi2_90) - 1)]
    if force_inliers is not None:
        inl_s = np.asarray(force_inliers[1], dtype=bool)
    tm["slam_rows"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # ---- Kalman update on the compressed measurement [H_slam ; R]
    Hc = np.zeros((2 * int(inl_s.sum()) + Rg.shape[0], N))
    Hc[:2 * int(inl_s.sum())] = Hs[inl_s].reshape(-1, N)
    Hc[2 * int(inl_s.sum()):, K_CORE:K_CORE + n] = Rg
    rc = np.concatenate([rs[inl_s].reshape(-1), zg])
    PHt = P @ Hc.T
    Sm = Hc @ PHt
    Sm[np.arange(len(Sm)), np.arange(len(Sm))] += var
    # the reference's covariance is not symmetric between updates, hence neither is S: LU, as its explicit inverse
    HP = Hc @ P
    Si = np.linalg.solve(Sm, np.concatenate([HP, rc[:, None]], axis=1))      # S^-1 [H P | r]
    delta = PHt @ Si[:, -1]
    Pn = P - PHt @ Si[:, :-1]                                                # (I - K H) P
    Pn = 0.5 * (Pn + Pn.T)
    tm["kalman_update"] = time.perf_counter() - t0
    return delta, {"seconds": tm, "gamma": gamma, "inlier": inl, "slam_inlier": inl_s, "rows": len(rc), "P_trace": float(np.trace(Pn))}
