"""fp64 contraction kernels at the cfg-5 shapes: TMA-staged tiles (k_gemm_tma.cu) vs the cp.async tiles (k_linalg.cu).
Run on a B200: python tools/gemm_bench.py   (device time per launch, achieved fp64 TFLOP/s against the 36.9 measured peak)"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from x_multi_agent_b200 import lib as L  # noqa: E402

lib = L.load()
rng = np.random.default_rng(0)
for name, M, N, K in (("schur cfg-5", 3136, 320, 1600), ("schur cfg-2", 1120, 192, 416), ("wide", 4096, 2048, 512)):
    A, B, Cm = rng.normal(size=(M, K)), rng.normal(size=(N, K)), rng.normal(size=(M, N))
    for op, tag in ((0, "tma"), (1, "cp.async")):
        ms = C.c_double(0.0)
        out = Cm.copy()
        used = lib.xb_debug_gemm(op, M, N, K, L.dptr(A), K, L.dptr(B), K, -1.0, 1.0, L.dptr(out), N, 20, C.cast(C.byref(ms), L.c_double_p))
        print(f"{name:12s} {M}x{N}x{K} {tag:9s} tma={used} {ms.value * 1e3:8.1f} us  {2.0 * M * N * K / ms.value / 1e9:6.2f} TFLOP/s")
for n, K in ((2715, 1600), (2715, 320), (795, 416)):
    W, P = rng.normal(size=(n, K)), rng.normal(size=(n, n))
    ms = C.c_double(0.0)
    used = lib.xb_debug_gemm(2, n, n, K, L.dptr(W), K, L.dptr(W), K, 0.0, 0.0, L.dptr(P), n, 20, C.cast(C.byref(ms), L.c_double_p))
    print(f"downdate n={n} K={K} tma={used} {ms.value * 1e3:8.1f} us  {1.0 * n * n * K / ms.value / 1e9:6.2f} TFLOP/s (useful, lower triangle)")
