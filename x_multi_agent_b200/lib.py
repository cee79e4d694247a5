"""ctypes binding of libxb200.so (the C ABI in include/xb200.h).

There is no CPU fallback: if the shared library is missing `load()` raises, and `xb_create` fails
when no sm_100 device is present.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["XB200_LIB"]) if os.environ.get("XB200_LIB") else _HERE / "libxb200.so"  # override: A/B builds

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class XbConfig(C.Structure):
    _fields_ = [
        ("n_poses_max", C.c_int), ("n_features_max", C.c_int), ("n_slots", C.c_int), ("n_generations", C.c_int),
        ("device", C.c_int), ("max_tracks", C.c_int), ("max_obs", C.c_int), ("iekf_iter", C.c_int),
        ("min_track_length", C.c_int), ("delta_seq_imu", C.c_uint), ("g", C.c_double * 3),
        ("n_w", C.c_double), ("n_bw", C.c_double), ("n_a", C.c_double), ("n_ba", C.c_double),
        ("a_m_max", C.c_double), ("time_margin", C.c_double),
        ("sigma_img", C.c_double), ("sigma_range", C.c_double), ("rho_0", C.c_double), ("sigma_rho_0", C.c_double),
        ("sigma_landmark", C.c_double), ("ci_msckf_w", C.c_double), ("ci_slam_w", C.c_double),
        ("downdate_precision", C.c_int), ("multi_uav", C.c_int), ("oc_projection", C.c_int),
    ]


class XbTrackList(C.Structure):
    _fields_ = [("n_tracks", C.c_int), ("off", c_int_p), ("obs", c_double_p)]


class XbMeasurement(C.Structure):
    _fields_ = [("timestamp", C.c_double), ("slam", XbTrackList), ("msckf", XbTrackList), ("msckf_short", XbTrackList),
                ("new_slam_std", XbTrackList), ("new_msckf_slam", XbTrackList), ("n_lost", C.c_int),
                ("lost_slam_idxs", c_int_p)]


class XbRangeMeasurement(C.Structure):
    """xb_range_measurement (include/xb200.h)."""
    _fields_ = [("timestamp", C.c_double), ("range", C.c_double), ("img_pt_n", C.c_double * 2),
                ("n_tr_feat_ids", C.c_int), ("tr_feat_ids", C.c_int * 3)]


class XbSunAngleMeasurement(C.Structure):
    """xb_sun_angle_measurement (include/xb200.h)."""
    _fields_ = [("timestamp", C.c_double), ("x_angle", C.c_double), ("y_angle", C.c_double)]


class XbPeerState(C.Structure):
    _fields_ = [("n_poses_max", C.c_int), ("n_features_max", C.c_int), ("positions", c_double_p),
                ("orientations", c_double_p), ("features", c_double_p), ("anchor_idxs", c_int_p), ("cov", c_double_p),
                ("cov_layout", C.c_int), ("translation", C.c_double * 3)]


class XbSlamMatch(C.Structure):
    _fields_ = [("peer", C.c_int), ("current_feature_id", C.c_int), ("received_feature_id", C.c_int)]


class XbMsckfMatch(C.Structure):
    _fields_ = [("peer", C.c_int), ("which", C.c_int), ("id_current_track", C.c_int), ("n_obs", C.c_int),
                ("obs", c_double_p)]


# every exported symbol of include/xb200.h: name -> (restype, argtypes)
_VP = C.c_void_p

class XbTmConfig(C.Structure):
    """xb_tm_config (include/xb200.h)."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("s", C.c_double),
                ("img_width", C.c_uint), ("img_height", C.c_uint), ("min_baseline_x_n", C.c_double),
                ("min_baseline_y_n", C.c_double), ("n_tiles_h", C.c_uint), ("n_tiles_w", C.c_uint), ("multi_uav", C.c_int)]


SIGNATURES = {
    "xb_tm_create": (_VP, [C.POINTER(XbTmConfig)]),
    "xb_tm_destroy": (None, [_VP]),
    "xb_tm_clear": (None, [_VP]),
    "xb_tm_manage_tracks": (C.c_int, [_VP, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_int]),
    "xb_tm_list_size": (C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "xb_tm_get_list": (C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
    "xb_tm_lost_slam_idxs": (C.c_int, [_VP, C.POINTER(C.c_int), C.c_int]),
    "xb_tm_remove_persistent_track": (C.c_int, [_VP, C.c_uint]),
    "xb_tm_remove_new_persistent_tracks": (C.c_int, [_VP, C.POINTER(C.c_uint), C.c_int]),
    "xb_tm_set_opp_ids": (C.c_int, [_VP, C.POINTER(C.c_ulonglong), C.c_int]),
    "xb_tm_counts": (C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "xb_tm_feature_triangle_at_point": (C.c_int, [_VP, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    "xb_tm_normalize_point": (C.c_int, [_VP, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "xb_tm_delaunay_facet": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                       C.POINTER(C.c_int)]),
    "xb_debug_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, C.c_double,
                                C.c_double, c_double_p, C.c_int, C.c_int, c_double_p]),
    "xb_default_config": (None, [C.POINTER(XbConfig)]),
    "xb_create": (C.c_int, [C.POINTER(XbConfig), C.POINTER(_VP)]),
    "xb_destroy": (C.c_int, [_VP]),
    "xb_last_error": (C.c_char_p, []),
    "xb_version": (C.c_char_p, []),
    "xb_set_stream": (C.c_int, [_VP, _VP]),
    "xb_synchronize": (C.c_int, [_VP]),
    "xb_n_error_states": (C.c_int, [_VP]),
    "xb_xvec_len": (C.c_int, [_VP]),
    "xb_ekf_initialize_from_state": (C.c_int, [_VP, c_double_p, c_double_p, C.c_int]),
    "xb_ekf_process_imu": (C.c_int, [_VP, C.c_double, C.c_uint, c_double_p, c_double_p, c_double_p]),
    "xb_ekf_process_imu_batch": (C.c_int, [_VP, C.c_int, c_double_p, C.POINTER(C.c_uint), c_double_p, c_double_p, c_double_p]),
    "xb_vio_set_measurement": (C.c_int, [_VP, C.POINTER(XbMeasurement)]),
    "xb_vio_set_sensors": (C.c_int, [_VP, C.POINTER(XbRangeMeasurement), C.POINTER(XbSunAngleMeasurement)]),
    "xb_host_alloc": (_VP, [C.c_size_t]),
    "xb_host_free": (None, [_VP]),
    "xb_ekf_process_update": (C.c_int, [_VP, c_double_p]),
    "xb_ekf_process_others": (C.c_int, [_VP, C.c_double, C.POINTER(XbPeerState), C.c_int, C.POINTER(XbSlamMatch),
                                        C.c_int, c_double_p]),
    "xb_vio_set_msckf_matches": (C.c_int, [_VP, C.POINTER(XbPeerState), C.c_int, C.POINTER(XbMsckfMatch), C.c_int]),
    "xb_vio_set_msckf_matches_packed": (C.c_int, [_VP, _VP, C.c_int, C.POINTER(XbMsckfMatch), C.c_int]),
    "xb_ci_pose_payload_len": (C.c_int, [_VP]),
    "xb_ci_pack_poses": (C.c_int, [_VP, C.c_int, _VP]),
    "xb_mm_last_gates": (C.c_int, [_VP, C.c_int, c_double_p, C.c_int]),
    "xb_updater_apply_ci_lists": (C.c_int, [_VP]),
    "xb_ekf_get_state": (C.c_int, [_VP, C.c_int, c_double_p]),
    "xb_ekf_get_covariance": (C.c_int, [_VP, C.c_int, c_double_p, C.c_int]),
    "xb_ekf_newest_slot": (C.c_int, [_VP]),
    "xb_sm_n_poses": (C.c_int, [_VP]),
    "xb_sm_n_features": (C.c_int, [_VP]),
    "xb_sm_anchor_idxs": (C.c_int, [_VP, c_int_p]),
    "xb_sm_set": (C.c_int, [_VP, C.c_int, C.c_int, c_int_p, C.c_int]),
    "xb_work_load": (C.c_int, [_VP, C.c_int]),
    "xb_work_store": (C.c_int, [_VP, C.c_int]),
    "xb_work_set": (C.c_int, [_VP, c_double_p, c_double_p, C.c_int]),
    "xb_work_get": (C.c_int, [_VP, c_double_p, c_double_p, C.c_int]),
    "xb_sm_manage": (C.c_int, [_VP, c_int_p, C.c_int]),
    "xb_vio_construct_update": (C.c_int, [_VP, C.c_int]),
    "xb_updater_reset_correction": (C.c_int, [_VP]),
    "xb_updater_apply_constructed": (C.c_int, [_VP, C.c_int]),
    "xb_updater_apply_update": (C.c_int, [_VP, c_double_p, c_double_p, c_double_p, C.c_int, c_double_p, C.c_int]),
    "xb_updater_apply_ci": (C.c_int, [_VP, c_double_p, c_double_p, c_double_p, C.c_int, c_int_p, C.c_int, C.c_double]),
    "xb_vio_post_update": (C.c_int, [_VP]),
    "xb_updater_update": (C.c_int, [_VP]),
    "xb_propagate": (C.c_int, [_VP, C.c_int, C.c_int]),
    "xb_ci_payload_len": (C.c_int, [_VP]),
    "xb_ci_pack": (C.c_int, [_VP, C.c_int, _VP]),
    "xb_ekf_process_others_packed": (C.c_int, [_VP, C.c_double, _VP, C.c_int, C.POINTER(XbSlamMatch), C.c_int, c_double_p]),
    "xb_ci_last_gates": (C.c_int, [_VP, c_double_p, C.c_int]),
    "xb_debug_read": (C.c_int, [_VP, C.c_char_p, c_double_p, C.c_int]),
    "xb_debug_read_int": (C.c_int, [_VP, C.c_char_p, c_int_p, C.c_int]),
    "xb_profile_enable": (C.c_int, [_VP, C.c_int]),
    "xb_profile_read": (C.c_int, [_VP, C.POINTER(C.c_char_p), c_double_p, C.POINTER(C.c_longlong), C.c_int]),
    "xb_kernel_launches": (C.c_longlong, [_VP]),
    "xb_chi2_quantile": (C.c_double, [C.c_double, C.c_double]),
    "xb_updater_collaborative_update": (C.c_int, [_VP, C.POINTER(XbPeerState), C.c_int, C.POINTER(XbSlamMatch), C.c_int]),
    "xb_ekf_update_begin": (C.c_int, [_VP, C.c_double, c_double_p]),
    "xb_ekf_update_end": (C.c_int, [_VP, c_double_p]),
    "xb_ekf_slot_info": (C.c_int, [_VP, C.c_int, c_double_p, c_int_p]),
    "xb_ekf_last_update_slot": (C.c_int, [_VP]),
}

_lib = None


def load():
    """Load libxb200.so and declare every signature.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(x_multi_agent_b200 has no CPU fallback)")
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("XB200_LIB") and not hasattr(lib, name):
            continue   # an older A/B build may lack newer entry points
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class XbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"xb200 error {code}: {msg}")
        self.code = code


def check(rc):
    if rc < 0:
        raise XbError(rc, load().xb_last_error().decode())
    return rc


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(c_int_p) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
