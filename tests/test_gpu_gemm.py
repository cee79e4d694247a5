"""The TMA-staged fp64 tensor-core contraction (k_gemm_tma.cu) against numpy, at the shapes the cfg-5 update produces and at
ragged shapes that exercise the zero-filled tile borders."""
import ctypes as C

import numpy as np
import pytest

from x_multi_agent_b200 import lib as L

pytestmark = pytest.mark.gpu


def _gemm(op, A, B, Cm, alpha, beta, reps=1):
    lib = L.load()
    A = np.ascontiguousarray(A)
    B = np.ascontiguousarray(B) if B is not None else np.zeros((1, A.shape[1]))
    Cm = np.ascontiguousarray(Cm).copy()
    ms = C.c_double(0.0)
    M, K = A.shape
    N = B.shape[0] if op != 2 else M
    used = lib.xb_debug_gemm(op, M, N, K, L.dptr(A), A.shape[1], L.dptr(B), B.shape[1], alpha, beta, L.dptr(Cm), Cm.shape[1], reps,
                             C.cast(C.byref(ms), L.c_double_p))
    assert used >= 0, lib.xb_last_error()
    return Cm, used, ms.value


@pytest.mark.parametrize("M,N,K", [(3136, 320, 1600), (1000, 250, 130), (2715 + 96, 333, 72), (4096, 200, 64), (12288, 64, 64)])
def test_tma_gemm_matches_numpy(M, N, K):
    rng = np.random.default_rng(M + N + K)
    lda = K + (K % 2)          # TMA: 16-byte aligned rows
    A = np.zeros((M, lda)); A[:, :K] = rng.normal(size=(M, K))
    B = np.zeros((N, lda)); B[:, :K] = rng.normal(size=(N, K))
    C0 = rng.normal(size=(M, N + 3))   # odd leading dimension on the output side (like P) is allowed
    want = C0.copy()
    want[:, :N] = 0.5 * C0[:, :N] - 1.25 * A[:, :K] @ B[:, :K].T
    got, used, _ = _gemm(0, A[:, :lda], B[:, :lda], C0, -1.25, 0.5)
    # the debug entry passes K = lda; the padding columns are zero
    big = ((M + 127) // 128) * ((N + 63) // 64) >= 96    # smaller problems stay on the cp.async tiles (launch latency)
    assert used == int(big), "unexpected kernel choice"
    assert np.abs(got[:, :N] - want[:, :N]).max() < 1e-11 * K
    assert np.array_equal(got[:, N:], C0[:, N:])
    ref, used1, _ = _gemm(1, A[:, :lda], B[:, :lda], C0, -1.25, 0.5)
    assert used1 == 0 and np.abs(ref[:, :N] - want[:, :N]).max() < 1e-11 * K


@pytest.mark.parametrize("n,K", [(2715, 1600), (2715, 320), (1301, 100)])
def test_tma_symmetric_downdate_matches_numpy(n, K):
    rng = np.random.default_rng(n + K)
    ldw = K + (K % 2)
    W = np.zeros((n, ldw)); W[:, :K] = rng.normal(size=(n, K)) * 0.1
    P = rng.normal(size=(n, n))     # deliberately unsymmetric: the kernel symmetrises like updater.cpp:131-136
    want = 0.5 * (P + P.T) - W @ W.T
    got, used, _ = _gemm(2, W, None, P, 0.0, 0.0)
    assert used == 1
    assert np.abs(got - want).max() < 1e-11 * K
    assert np.array_equal(got, got.T)
