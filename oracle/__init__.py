"""fp64 CPU oracle for the xVIO EKF/MSCKF hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference filter arithmetic (jpl-x/x_multi_agent,
`/root/reference`); every function cites the reference file:line it follows.  It is the checker
for the CUDA path, never the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product package
(`x_multi_agent_b200`) never imports it and fails loudly when its CUDA library is missing.

PARITY PINNING: pinned against the reference itself, run in this container.  The reference ships no tests,
golden vectors or fixtures, and Eigen / OpenCV C++ / Boost / NLopt are not installed, so
`oracle/ref_build/build_ref.sh` compiles the reference's UNMODIFIED filter back end where it lies
(/root/reference/src/x/{ekf,vio,vision}/*.cpp, both the single-agent and the -DMULTI_UAV flavour) against
stand-in headers written from scratch (`oracle/ref_build/shim/`: a small eager Eigen, cv::Mat + the two-view DLT,
the chi-squared quantile, a no-op logger, an NLopt whose optimiser always fails) into `oracle/_ref/libxref*.so`;
`oracle/refcpp.py` binds it.  `tests/test_ref_pinning.py` holds this package to that binary's outputs:
whole sequences through Ekf::processImu / processUpdateMeasurement / processOthersMeasurement (state 1e-10,
covariance 1e-10 relative -- observed 1e-14 --, incl. the unsymmetrised covariance, short tracks, MSCKF-SLAM
promotion, IEKF, the multi-agent MSCKF block and SLAM-SLAM CI) and the stage methods (applyUpdate,
applyQRDecomposition, StateManager::manage, Propagator, MsckfUpdate); `tests/golden/ref_sequences.npz`
(written by `oracle/tools/make_ref_golden.py` from the same binary) carries the pin to machines without
/root/reference.  What the stand-ins replace is third-party arithmetic, not reference code: products / LU /
Householder QR are restated with the same conventions as Eigen's (differences are summation order only), and
`cv::triangulatePoints` only seeds the Gauss-Newton refinement.  `qd_poly` is additionally pinned against
`propagator.cpp:207-840` compiled with no stand-in at all (`oracle/_ref/libxref_qd.so`).
Not reproducible by the stand-in build: NLopt-optimised CI weights (ci.cpp:129-190, time-limited COBYLA).
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
