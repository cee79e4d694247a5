"""ctypes binding of the reference's OWN filter back end compiled in place (oracle/_ref/libxref*.so) -- TEST
INFRASTRUCTURE ONLY.

`oracle/ref_build/build_ref.sh` compiles the unmodified sources under /root/reference/src/x/{ekf,vio,vision}
against stand-in headers for Eigen / OpenCV / Boost / NLopt (oracle/ref_build/shim/, written from scratch; none of the
four is installed here) plus `xref_harness.cpp`, which plays x::VIO.  `RefFilter` gives that binary the call surface of
`x_multi_agent_b200.Filter` / `tests/oracle_driver.OracleFilter`, so one recorded event stream can be replayed on the
CUDA path, on the numpy oracle and on the reference itself.  The built libraries travel to the GPU box; the sources
do not (nothing here reads /root/reference at run time).
"""
import ctypes as C
import glob
import os
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent / "_ref"
_LIBS = {}
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_ulonglong)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def lib_path(flavour="single"):
    name = {"single": "libxref.so", "multi": "libxref_multi.so", "release": "libxref_release.so"}[flavour]
    return _DIR / name


def available(flavour="single"):
    return lib_path(flavour).exists()


def load(flavour="single"):
    if flavour in _LIBS:
        return _LIBS[flavour]
    p = lib_path(flavour)
    if not p.exists():
        raise FileNotFoundError(f"{p} missing: run oracle/ref_build/build_ref.sh where /root/reference exists")
    lib = C.CDLL(os.fspath(p))
    lib.xref_create.restype = C.c_void_p
    lib.xref_create.argtypes = [_dp]
    lib.xref_destroy.argtypes = [C.c_void_p]
    lib.xref_set_blas.argtypes = [C.c_char_p, C.c_int]
    lib.xref_init.argtypes = [C.c_void_p, _dp, _dp]
    lib.xref_process_imu.argtypes = [C.c_void_p, C.c_double, C.c_uint, _dp, _dp, _dp]
    lib.xref_set_measurement.argtypes = [C.c_void_p, C.c_double, _ip, C.POINTER(_ip), C.POINTER(_dp), C.POINTER(_up),
                                         _ip, C.c_int]
    lib.xref_process_update.argtypes = [C.c_void_p, _dp, _dp]
    lib.xref_sensor_rows.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_double, C.c_double, _ip, C.c_double,
                                     C.c_double, _dp, _dp, _dp]
    lib.xref_construct_update.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_int]
    if hasattr(lib, "xref_set_sensors"):
        lib.xref_set_sensors.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, _ip, C.c_int,
                                         C.c_double, C.c_double, C.c_double]
    lib.xref_get_state.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    lib.xref_sm_info.argtypes = [C.c_void_p, _ip, _ip, _ip]
    lib.xref_apply_update.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int]
    lib.xref_qr_compress.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, _dp, _dp]
    lib.xref_manage.argtypes = [C.c_void_p, _dp, _dp, _ip, C.c_int]
    lib.xref_propagate.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp, _dp, _dp, _dp]
    lib.xref_sm_set.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, C.c_int]
    lib.xref_updater_update.argtypes = [C.c_void_p, _dp, _dp, _dp]
    lib.xref_msckf_rows.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int, _ip, _dp, _dp, _dp, C.c_int]
    _LIBS[flavour] = lib
    return lib


def openblas_path():
    """The ILP64 OpenBLAS that ships with numpy (symbols scipy_*_64_)."""
    hits = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
    return os.path.realpath(hits[0]) if hits else None


def bind_blas(flavour="single", threads=0):
    """Route the stand-in's large GEMM / QR / LU to OpenBLAS (threads <= 0: library default).  Returns True if bound."""
    p = openblas_path()
    return bool(p) and load(flavour).xref_set_blas(p.encode(), int(threads)) == 0


def _csr(tracks):
    off = np.zeros(len(tracks) + 1, dtype=np.int32)
    if len(tracks):
        off[1:] = np.cumsum([np.asarray(t).reshape(-1, 2).shape[0] for t in tracks])
        obs = np.ascontiguousarray(np.vstack([np.asarray(t, dtype=np.float64).reshape(-1, 2) for t in tracks]))
    else:
        obs = np.zeros((1, 2))
    return off, obs


class RefState:
    """Estimates + covariance of one x::State in the xvec layout of include/xb200.h."""

    def __init__(self, M, F, x, cov=None):
        self.M, self.F, self.x, self.cov = M, F, x, cov

    p = property(lambda s: s.x[0:3])
    v = property(lambda s: s.x[3:6])
    q = property(lambda s: s.x[6:10])
    b_w = property(lambda s: s.x[10:13])
    b_a = property(lambda s: s.x[13:16])
    time = property(lambda s: float(s.x[29]))
    p_array = property(lambda s: s.x[32:32 + 3 * s.M])
    q_array = property(lambda s: s.x[32 + 3 * s.M:32 + 7 * s.M])
    f_array = property(lambda s: s.x[32 + 7 * s.M:32 + 7 * s.M + 3 * s.F])

    def copy(self):
        return RefState(self.M, self.F, self.x.copy(), None if self.cov is None else self.cov.copy())


class RefFilter:
    """x::Ekf + x::VioUpdater + x::StateManager of the compiled reference (src/x/ekf/ekf.cpp, src/x/vio/vio_updater.cpp,
    src/x/vio/state_manager.cpp), built as VIO::setUp does (vio.cpp:176-214)."""

    def __init__(self, M, F, sigma_img=1.0 / 320.0, rho_0=0.5, sigma_rho_0=0.25, iekf_iter=1, n_slots=250,
                 g=(0.0, 0.0, -9.81), noise=None, flavour="single", sigma_landmark=0.0, ci_msckf_w=-1.0,
                 ci_slam_w=-1.0, a_m_max=50.0, time_margin=0.005, min_track_length=0, sigma_range=0.0):
        self.lib = load(flavour)
        self.M, self.F = M, F
        self.N = 15 + 6 * M + 3 * F
        self.LX = 32 + 7 * M + 3 * F
        nz = noise if noise is not None else (0.0083, 0.00083, 0.0013, 0.00013)
        if not isinstance(nz, (tuple, list)):
            nz = (nz.n_w, nz.n_bw, nz.n_a, nz.n_ba)
        cfg = np.array([M, F, n_slots, sigma_img, sigma_range, rho_0, sigma_rho_0, min_track_length, sigma_landmark,
                        ci_msckf_w, ci_slam_w, iekf_iter, *g, *nz, a_m_max, time_margin], dtype=np.float64)
        self.h = C.c_void_p(self.lib.xref_create(_d(cfg)))
        if not self.h:
            raise RuntimeError("xref_create failed")
        self.last_seconds = 0.0
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.xref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- x::Ekf ------------------------------------------------------------------------------------------
    def initialize_from_state(self, xs):
        x = np.ascontiguousarray(xs.x, dtype=np.float64)
        cov = np.ascontiguousarray(xs.cov, dtype=np.float64)
        rc = self.lib.xref_init(self.h, _d(x), _d(cov))
        if rc == -2:
            raise ValueError("init_bfr_mismatch")
        if rc:
            raise RuntimeError("Ekf::initializeFromState failed")

    def process_imu(self, t, seq, w_m, a_m, want_state=True):
        out = np.empty(self.LX) if want_state else None
        w = np.ascontiguousarray(w_m, dtype=np.float64)
        a = np.ascontiguousarray(a_m, dtype=np.float64)
        rc = self.lib.xref_process_imu(self.h, float(t), int(seq), _d(w), _d(a), _d(out))
        if rc == 0:
            return None
        return RefState(self.M, self.F, out) if want_state else True

    def set_measurement(self, m, ids=None):
        lists = (m.slam_trks, m.msckf_trks, m.msckf_short_trks, m.new_slam_std_trks, m.new_msckf_slam_trks)
        n = np.array([len(t) for t in lists], dtype=np.int32)
        csr = [_csr(t) for t in lists]
        offs = (_ip * 5)(*[c[0].ctypes.data_as(_ip) for c in csr])
        obs = (_dp * 5)(*[c[1].ctypes.data_as(_dp) for c in csr])
        idp = None
        id_arrs = None
        if ids is not None:
            id_arrs = [np.ascontiguousarray(ids.get(k, np.arange(len(t)) + 1 + 1000000 * k), dtype=np.uint64)
                       for k, t in enumerate(lists)]
            idp = (_up * 5)(*[a.ctypes.data_as(_up) for a in id_arrs])
        lost = np.ascontiguousarray(m.lost_slam_trk_idxs, dtype=np.int32)
        self._keep = (n, csr, offs, obs, id_arrs, idp, lost)
        self.lib.xref_set_measurement(self.h, float(m.timestamp), n.ctypes.data_as(_ip), offs, obs, idp,
                                      lost.ctypes.data_as(_ip), len(lost))
        rng, sun = getattr(m, "range", None), getattr(m, "sun_angle", None)
        if rng is not None or sun is not None:  # VioMeasurement::range / sun_angle (vio/types.h:300-305)
            tri = np.ascontiguousarray(list(rng.tr_feat_ids) if rng is not None else [], dtype=np.int32)
            self.lib.xref_set_sensors(self.h, float(rng.timestamp) if rng is not None else -1.0,
                                      float(rng.range) if rng is not None else 0.0,
                                      float(rng.img_pt_n[0]) if rng is not None else 0.0,
                                      float(rng.img_pt_n[1]) if rng is not None else 0.0,
                                      tri.ctypes.data_as(_ip), len(tri),
                                      float(sun.timestamp) if sun is not None else -1.0,
                                      float(sun.x_angle) if sun is not None else 0.0,
                                      float(sun.y_angle) if sun is not None else 0.0)

    def process_update_measurement(self, want_state=True):
        out = np.empty(self.LX)
        sec = C.c_double(0.0)
        rc = self.lib.xref_process_update(self.h, _d(out), C.cast(C.byref(sec), _dp))
        self.last_seconds = sec.value
        if rc == 0:
            return None
        return RefState(self.M, self.F, out) if want_state else True

    def get_state(self, which=-1, with_cov=True):
        x = np.empty(self.LX)
        cov = np.empty((self.N, self.N)) if with_cov else None
        if self.lib.xref_get_state(self.h, int(which), _d(x), _d(cov)):
            raise RuntimeError("xref_get_state failed")
        return RefState(self.M, self.F, x, cov)

    def newest(self):
        return self.get_state(-1)

    def get_covariance(self, which=-1):
        return self.get_state(which).cov

    def sm_info(self):
        n_p, n_f = C.c_int(0), C.c_int(0)
        an = np.zeros(max(self.F, 1), dtype=np.int32)
        self.lib.xref_sm_info(self.h, C.byref(n_p), C.byref(n_f), an.ctypes.data_as(_ip))
        return n_p.value, n_f.value, an[:self.F].tolist()

    def sm_set(self, n_poses, n_features, anchor_idxs, filled_before):
        a = np.full(max(self.F, 1), -1, dtype=np.int32)
        a[:len(anchor_idxs)] = anchor_idxs
        self.lib.xref_sm_set(self.h, int(n_poses), int(n_features), a.ctypes.data_as(_ip), int(filled_before))

    def updater_update(self, xs):
        """Updater::update (updater.cpp:39-115) on a copy of `xs` with the measurement set by set_measurement()."""
        x = np.ascontiguousarray(xs.x, dtype=np.float64).copy()
        cov = np.ascontiguousarray(xs.cov, dtype=np.float64).copy()
        sec = C.c_double(0.0)
        self.lib.xref_updater_update(self.h, _d(x), _d(cov), C.cast(C.byref(sec), _dp))
        self.last_seconds = sec.value
        return RefState(self.M, self.F, x, cov)

    # ---- stage level ---------------------------------------------------------------------------------------
    def apply_update(self, xs, H, res, r_diag, correction_total, cov_update=True):
        """Updater::applyUpdate (updater.cpp:117-141) on a copy of `xs`; returns (state, correction_total)."""
        x = np.ascontiguousarray(xs.x, dtype=np.float64).copy()
        cov = np.ascontiguousarray(xs.cov, dtype=np.float64).copy()
        H = np.ascontiguousarray(H, dtype=np.float64)
        res = np.ascontiguousarray(res, dtype=np.float64).ravel()
        rd = np.ascontiguousarray(r_diag, dtype=np.float64).ravel()
        ct = np.ascontiguousarray(correction_total, dtype=np.float64).ravel().copy()
        self.lib.xref_apply_update(self.h, _d(x), _d(cov), _d(H), _d(res), _d(rd), H.shape[0], _d(ct), int(cov_update))
        return RefState(self.M, self.F, x, cov), ct

    def qr_compress(self, H, res):
        """VioUpdater::applyQRDecomposition (vio_updater.cpp:487-512)."""
        H = np.ascontiguousarray(H, dtype=np.float64)
        res = np.ascontiguousarray(res, dtype=np.float64).ravel()
        m, n = H.shape
        Ho, ro = np.zeros((max(m, n), n)), np.zeros(max(m, n))
        mo = self.lib.xref_qr_compress(self.h, _d(H), _d(res), m, n, _d(Ho), _d(ro))
        return Ho[:mo].copy(), ro[:mo].copy()

    def manage(self, xs, lost=()):
        """StateManager::manage (state_manager.cpp:31-149) with this filter's bookkeeping."""
        x = np.ascontiguousarray(xs.x, dtype=np.float64).copy()
        cov = np.ascontiguousarray(xs.cov, dtype=np.float64).copy()
        lo = np.ascontiguousarray(list(lost), dtype=np.int32)
        self.lib.xref_manage(self.h, _d(x), _d(cov), lo.ctypes.data_as(_ip), len(lo))
        return RefState(self.M, self.F, x, cov)

    def propagate(self, xs, t1, w1, a1):
        """Propagator::propagateState + propagateCovariance (propagator.cpp:30-72)."""
        x0 = np.ascontiguousarray(xs.x, dtype=np.float64)
        c0 = np.ascontiguousarray(xs.cov, dtype=np.float64)
        w = np.ascontiguousarray(w1, dtype=np.float64)
        a = np.ascontiguousarray(a1, dtype=np.float64)
        x1, c1 = np.empty(self.LX), np.empty((self.N, self.N))
        self.lib.xref_propagate(self.h, _d(x0), _d(c0), float(t1), _d(w), _d(a), _d(x1), _d(c1))
        return RefState(self.M, self.F, x1, c1)

    def sensor_rows(self, xs, rng, sun):
        """RangeUpdate (range_update.cpp:24-265) + SolarUpdate (solar_update.cpp:25-94) on `xs`: (jac 3 x N, res 3,
        r_diag 3), row 0 the range row, rows 1-2 the sun-sensor rows."""
        x = np.ascontiguousarray(xs.x, dtype=np.float64)
        cov = np.ascontiguousarray(xs.cov, dtype=np.float64)
        tri = np.ascontiguousarray(list(rng.tr_feat_ids), dtype=np.int32)
        J, r, d = np.zeros((3, self.N)), np.zeros(3), np.zeros(3)
        self.lib.xref_sensor_rows(self.h, _d(x), _d(cov), float(rng.range), float(rng.img_pt_n[0]), float(rng.img_pt_n[1]),
                                  tri.ctypes.data_as(_ip), float(sun.x_angle), float(sun.y_angle), _d(J), _d(r), _d(d))
        return J, r, d

    def construct_update(self, xs, max_rows=4096):
        """VioUpdater::constructUpdate (vio_updater.cpp:266-423, single-UAV build) for the measurement set with
        set_measurement() on the post-manage state `xs`: (h, res, diag R)."""
        x = np.ascontiguousarray(xs.x, dtype=np.float64)
        cov = np.ascontiguousarray(xs.cov, dtype=np.float64)
        H, r, d = np.zeros((max_rows, self.N)), np.zeros(max_rows), np.zeros(max_rows)
        rows = self.lib.xref_construct_update(self.h, _d(x), _d(cov), _d(H), _d(r), _d(d), max_rows)
        if rows < 0:
            raise RuntimeError(f"xref_construct_update failed ({rows})")
        return H[:rows].copy(), r[:rows].copy(), d[:rows].copy()

    def msckf_rows(self, xs, tracks, timestamp=0.0):
        """MsckfUpdate (msckf_update.cpp:27-63) on `xs`: stacked inlier Jacobian rows and residual."""
        x = np.ascontiguousarray(xs.x, dtype=np.float64)
        cov = np.ascontiguousarray(xs.cov, dtype=np.float64)
        off, obs = _csr(tracks)
        max_rows = int(2 * off[-1])
        J, r = np.zeros((max_rows, self.N)), np.zeros(max_rows)
        rows = self.lib.xref_msckf_rows(self.h, _d(x), _d(cov), float(timestamp), len(tracks), off.ctypes.data_as(_ip),
                                        _d(obs), _d(J), _d(r), max_rows)
        if rows < 0:
            raise RuntimeError(f"xref_msckf_rows failed ({rows})")
        return J[:rows].copy(), r[:rows].copy()


# ---- MULTI_UAV flavour (libxref_multi.so) ----------------------------------------------------------------------
class _XrefPeer(C.Structure):
    _fields_ = [("M", C.c_int), ("F", C.c_int), ("dynamic", _dp), ("positions", _dp), ("orientations", _dp),
                ("features", _dp), ("anchors", _ip), ("cov_rm", _dp)]


def _peers(peers, keep):
    """peers: objects with positions (3M), orientations (4M), features (3F), anchor_idxs, cov (x::SimpleState)."""
    arr = (_XrefPeer * max(1, len(peers)))()
    for i, p in enumerate(peers):
        pos = np.ascontiguousarray(p.positions, dtype=np.float64)
        ori = np.ascontiguousarray(p.orientations, dtype=np.float64)
        fe = np.ascontiguousarray(p.features, dtype=np.float64)
        cov = np.ascontiguousarray(p.cov, dtype=np.float64)
        an = np.ascontiguousarray(list(p.anchor_idxs) + [0], dtype=np.int32)
        dyn = np.zeros(16)
        keep += [pos, ori, fe, cov, an, dyn]
        arr[i].M, arr[i].F = pos.size // 3, fe.size // 3
        arr[i].dynamic, arr[i].positions, arr[i].orientations = _d(dyn), _d(pos), _d(ori)
        arr[i].features, arr[i].anchors, arr[i].cov_rm = _d(fe), an.ctypes.data_as(_ip), _d(cov)
    return arr


MSCKF_ID0, SHORT_ID0 = 1000001, 2000001   # Track ids RefFilter.set_measurement(ids={}) gives the two MSCKF lists


class RefFilterMulti(RefFilter):
    """The -DMULTI_UAV build of the reference (Updater::update with CI lists, Ekf::processOthersMeasurement)."""

    def __init__(self, M, F, **kw):
        kw.setdefault("flavour", "multi")
        super().__init__(M, F, **kw)
        lib = self.lib
        lib.xref_set_msckf_matches.argtypes = [C.c_void_p, C.POINTER(_XrefPeer), C.c_int, C.c_int, _ip, _up, _ip, _dp,
                                               C.c_double]
        lib.xref_process_others.argtypes = [C.c_void_p, C.c_double, C.POINTER(_XrefPeer), C.c_int, C.c_int, _ip, _ip,
                                            _ip, _dp]

    def set_measurement(self, m, ids=None):
        super().set_measurement(m, ids={} if ids is None else ids)

    def set_msckf_matches(self, peers, matches, timestamp=0.0):
        """matches: (peer_idx, which, own_track_idx, received_track (L,2)) in list order (vision/types.h:83-100);
        which = 0 -> msckf_trks, 1 -> msckf_short_trks of the measurement set with set_measurement()."""
        keep = []
        cp = _peers(peers, keep)
        peer_of = np.ascontiguousarray([m[0] for m in matches], dtype=np.int32)
        own = np.ascontiguousarray([(MSCKF_ID0 if m[1] == 0 else SHORT_ID0) + m[2] for m in matches], dtype=np.uint64)
        off, obs = _csr([m[3] for m in matches])
        self.lib.xref_set_msckf_matches(self.h, cp, len(peers), len(matches), peer_of.ctypes.data_as(_ip),
                                        own.ctypes.data_as(_up), off.ctypes.data_as(_ip), _d(obs), float(timestamp))

    def process_others_measurement(self, t, peers, matches, want_state=True):
        """Ekf::processOthersMeasurement (ekf.cpp:143-176); matches: (peer_idx, current_feature_id, received_id)."""
        keep = []
        cp = _peers(peers, keep)
        pe = np.ascontiguousarray([m[0] for m in matches], dtype=np.int32)
        cu = np.ascontiguousarray([m[1] for m in matches], dtype=np.int32)
        rc_ = np.ascontiguousarray([m[2] for m in matches], dtype=np.int32)
        out = np.empty(self.LX)
        rc = self.lib.xref_process_others(self.h, float(t), cp, len(peers), len(matches), pe.ctypes.data_as(_ip),
                                          cu.ctypes.data_as(_ip), rc_.ctypes.data_as(_ip), _d(out))
        if rc < 0:
            raise RuntimeError("Ekf::processOthersMeasurement threw")
        if rc == 0:
            return None
        return RefState(self.M, self.F, out) if want_state else True
