"""x_multi_agent_b200 -- B200-native (sm_100a) EKF/MSCKF visual-inertial update hot path of the X library.

The package holds only what the hot path needs: `csrc/` (hand-written CUDA kernels + the C ABI of
include/xb200.h, built into libxb200.so), `lib` (ctypes binding), `filter` (host-side mirror of the
reference's x::Ekf / x::VioUpdater / x::State API), `track_manager` (mirror of x::TrackManager: matches in, the
five track lists out) and `synth` (seeded synthetic inputs at the
VioUpdater::preProcess seam).  There is no CPU fallback.
"""
from .filter import Filter, Measurement, PackedMeasurement, PeerState, State  # noqa: F401
from .lib import LIB_PATH, XbError, load  # noqa: F401
from .track_manager import TrackManager  # noqa: F401
from .vio import VIO, load_params_from_yaml  # noqa: F401
