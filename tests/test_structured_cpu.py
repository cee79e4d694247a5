"""oracle/structured.py (the structured formulation on the CPU, bench.py's `cpu_baseline.structured`) against the dense
restatement of the reference (oracle.updates: MsckfUpdate / SlamUpdate, dense Kalman gain) on the same inputs: identical
gates and the same state correction.  This is what licenses reading its timing as "the device's algorithm on the host"."""
import numpy as np

import oracle.updates as upd
from oracle.structured import structured_update
from oracle.updates import MsckfUpdate, SlamUpdate, chi2_quantile


def _scene(seed, M, F, K):
    rng = np.random.default_rng(seed)
    ps = np.stack([np.linspace(0, 2.0, M), 0.1 * rng.standard_normal(M), 0.05 * rng.standard_normal(M)], axis=1)
    qs = np.concatenate([0.03 * rng.standard_normal((M, 3)), np.ones((M, 1))], axis=1)
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    from oracle.quat import rot
    Rs = [rot(q) for q in qs]
    lm = np.stack([rng.uniform(-2, 4, K + F), rng.uniform(-2, 2, K + F), rng.uniform(5, 10, K + F)], axis=1)

    def obs(pt, i):
        c = Rs[i].T @ (pt - ps[i])
        return c[:2] / c[2] + 2e-3 * rng.standard_normal(2)
    Z = np.array([[obs(lm[k], i) for i in range(M)] for k in range(K)])
    anchors = rng.integers(0, M - 1, size=F)
    fs = np.zeros((F, 3))
    slam_obs = np.zeros((F, 2))
    for j in range(F):
        c = Rs[anchors[j]].T @ (lm[K + j] - ps[anchors[j]])
        fs[j] = [c[0] / c[2], c[1] / c[2], 1.0 / c[2]]
        fs[j] += [1e-3 * rng.standard_normal(), 1e-3 * rng.standard_normal(), 2e-3 * rng.standard_normal()]
        slam_obs[j] = obs(lm[K + j], M - 1)
    N = 15 + 6 * M + 3 * F
    A = rng.standard_normal((N, N))
    P = 1e-4 * (A @ A.T / N + 0.5 * np.eye(N))
    return ps, qs, fs, anchors, P, Z, slam_obs


def test_structured_formulation_equals_the_dense_one():
    M, F, K, sigma = 6, 5, 9, 2e-3
    ps, qs, fs, anchors, P, Z, slam_obs = _scene(4, M, F, K)
    slam_len = np.full(F, 4)
    chi95 = np.array([0.0] + [chi2_quantile(0.95, d) for d in range(1, 2 * M + 1)])
    chi90 = np.array([0.0] + [chi2_quantile(0.90, d) for d in range(1, 64)])
    delta, info = structured_update(ps.reshape(-1), qs.reshape(-1), fs.reshape(-1), list(anchors), P, Z, slam_obs, slam_len, M, M, F,
                                    sigma, chi95, chi90)
    old = upd.OC_PROJECTION
    upd.OC_PROJECTION = False
    try:
        ms = MsckfUpdate([z for z in Z], list(qs), list(ps), P, M, sigma)
    finally:
        upd.OC_PROJECTION = old
    sl = SlamUpdate([np.vstack([np.zeros((3, 2)), slam_obs[j][None]]) for j in range(F)], list(qs), list(ps), fs.reshape(-1),
                    list(anchors), P, M, sigma)
    assert np.array_equal(ms.inlier, info["inlier"]) and ms.inlier.sum() >= K - 2
    assert np.allclose(ms.gamma, info["gamma"], rtol=1e-7)
    assert np.array_equal(sl.inlier, info["slam_inlier"])
    H = np.vstack([ms.jac[:ms.rows_used], sl.jac[:sl.rows_used]])
    r = np.concatenate([ms.res[:ms.rows_used], sl.res[:sl.rows_used]])
    S = H @ P @ H.T + sigma * sigma * np.eye(len(r))
    d_dense = P @ H.T @ np.linalg.solve(S, r)
    assert np.linalg.norm(delta - d_dense) < 1e-7 * np.linalg.norm(d_dense), np.linalg.norm(delta - d_dense) / np.linalg.norm(d_dense)
