#!/usr/bin/env python
"""Generates tests/golden/qd_reference.npz: inputs and outputs of the reference's own
Propagator::discreteProcessNoiseCov (reference: src/x/ekf/propagator.cpp:207-840), evaluated by the function
compiled where it lies (oracle/_ref/libxref_qd.so, recipe oracle/ref_build/build_ref.sh).  Run in the build
container only (needs /root/reference); the fixture travels, the reference does not."""
import ctypes
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
lib = ctypes.CDLL(str(ROOT / "oracle" / "_ref" / "libxref_qd.so"))
P = ctypes.POINTER(ctypes.c_double)
lib.xref_qd.argtypes = [ctypes.c_double, P, P, P] + [ctypes.c_double] * 4 + [P]
rng = np.random.default_rng(20260101)
n = 64
dt = rng.uniform(1e-3, 0.1, n)
q = rng.normal(size=(n, 4))
q /= np.linalg.norm(q, axis=1, keepdims=True)          # (x,y,z,w)
w = rng.normal(size=(n, 3)) * 0.7
a = rng.normal(size=(n, 3)) * 6.0
noise = np.column_stack([rng.uniform(1e-3, 1e-2, n), rng.uniform(1e-4, 1e-3, n), rng.uniform(1e-3, 1e-2, n),
                         rng.uniform(1e-5, 1e-3, n)])
out = np.zeros((n, 15, 15))
for i in range(n):
    qw = np.ascontiguousarray([q[i, 3], q[i, 0], q[i, 1], q[i, 2]])
    o = np.zeros(225)
    lib.xref_qd(dt[i], qw.ctypes.data_as(P), np.ascontiguousarray(w[i]).ctypes.data_as(P),
                np.ascontiguousarray(a[i]).ctypes.data_as(P), *noise[i], o.ctypes.data_as(P))
    out[i] = o.reshape(15, 15)
np.savez_compressed(ROOT / "tests" / "golden" / "qd_reference.npz", dt=dt, q=q, w=w, a=a, noise=noise, Qd=out)
print("wrote", ROOT / "tests" / "golden" / "qd_reference.npz", out.shape)
