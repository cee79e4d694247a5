"""fp64 CPU oracle for the xVIO EKF/MSCKF hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference filter arithmetic (jpl-x/x_multi_agent,
`/root/reference`); every function cites the reference file:line it follows.  It is the checker
for the CUDA path, never the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product package
(`x_multi_agent_b200`) never imports it and fails loudly when its CUDA library is missing.

PARITY PINNING.  The reference ships no tests, golden vectors or fixtures, and
its hot-path sources cannot be compiled here as a whole (Eigen, OpenCV C++, Boost, NLopt absent).
What *is* pinned against the reference's own code run in this container:
  * `qd_poly` (the 600-line symbolic process-noise polynomial) against
    `propagator.cpp:207-840` compiled where it lies into `oracle/_ref/libxref_qd.so`
    (recipe: `oracle/ref_build/build_ref.sh`).
Everything else is "parity unpinned": a line-by-line restatement with Eigen semantics
(Quaterniond(w,x,y,z) vs coeffs()=(x,y,z,w), toRotationMatrix, PartialPivLU inverse,
HouseholderQR) mirrored by numpy/LAPACK, plus basis-/sign-invariant checks.
"""
from .quat import rot, qmul, qnormalized, small_angle_quat, skew  # noqa: F401
from .state import State  # noqa: F401
from .propagator import Propagator, ImuNoise  # noqa: F401
from .state_manager import StateManager  # noqa: F401
from .triangulation import Triangulation  # noqa: F401
from .updates import MsckfUpdate, SlamUpdate, MsckfSlamUpdate, chi2_quantile  # noqa: F401
from .updater import VioUpdaterOracle, apply_update, apply_ci, apply_qr_decomposition  # noqa: F401
from .ci import fuse_ci_pair, fuse_ci_multi, MultiSlamUpdate, SimpleState  # noqa: F401
from .ekf import Ekf, StateBuffer  # noqa: F401
