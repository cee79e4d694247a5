"""Host-side Python mirror of the reference's operator API for the hot path, on top of the C ABI.

Names follow the reference: x::Ekf (include/x/ekf/ekf.h:53-195), x::VioUpdater
(include/x/vio/vio_updater.h:35-335), x::State (include/x/ekf/state.h:36-337), x::StateManager.
All arithmetic runs in libxb200.so on the GPU; this module only marshals buffers.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import lib as L

K_CORE = 15


def xvec_len(M, F):
    return 32 + 7 * M + 3 * F


class State:
    """Estimates of one x::State as views into a flat xvec (layout: include/xb200.h)."""

    def __init__(self, M, F, xvec=None):
        self.M, self.F = M, F
        self.x = np.zeros(xvec_len(M, F)) if xvec is None else np.array(xvec, dtype=np.float64)
        if xvec is None:
            self.x[9] = 1.0   # q = identity (x,y,z,w)
            self.x[19] = 1.0  # q_ic
            self.x[29] = -1.0  # kInvalid time (common/types.h:90)
        self.cov = None

    p = property(lambda s: s.x[0:3])
    v = property(lambda s: s.x[3:6])
    q = property(lambda s: s.x[6:10])
    b_w = property(lambda s: s.x[10:13])
    b_a = property(lambda s: s.x[13:16])
    q_ic = property(lambda s: s.x[16:20])
    p_ic = property(lambda s: s.x[20:23])
    w_m = property(lambda s: s.x[23:26])
    a_m = property(lambda s: s.x[26:29])
    p_array = property(lambda s: s.x[32:32 + 3 * s.M])
    q_array = property(lambda s: s.x[32 + 3 * s.M:32 + 7 * s.M])
    f_array = property(lambda s: s.x[32 + 7 * s.M:32 + 7 * s.M + 3 * s.F])

    @property
    def time(self):
        return float(self.x[29])

    @time.setter
    def time(self, t):
        self.x[29] = t

    def n_error_states(self):
        return K_CORE + 6 * self.M + 3 * self.F

    @classmethod
    def from_oracle(cls, s):
        """Build from an oracle.State (tests only)."""
        M, F = s.n_poses_max(), s.n_features_max()
        o = cls(M, F)
        o.x[0:3], o.x[3:6], o.x[6:10], o.x[10:13], o.x[13:16] = s.p, s.v, s.q, s.b_w, s.b_a
        o.x[16:20], o.x[20:23], o.x[23:26], o.x[26:29] = s.q_ic, s.p_ic, s.w_m, s.a_m
        o.x[29], o.x[30] = s.time, s.seq
        o.p_array[:], o.q_array[:], o.f_array[:] = s.p_array, s.q_array, s.f_array
        o.cov = np.array(s.cov, dtype=np.float64)
        return o


@dataclass
class RangeMeasurement:
    """x::RangeMeasurement (include/x/vio/types.h:223-243) + the SLAM-feature facet the LRF beam hits
    (TrackManager::featureTriangleAtPoint, vio_updater.cpp:365-366).  Used when timestamp > 0.1."""
    timestamp: float = -1.0
    range: float = 0.0
    img_pt_n: tuple = (0.0, 0.0)
    tr_feat_ids: list = field(default_factory=list)


@dataclass
class SunAngleMeasurement:
    """x::SunAngleMeasurement (include/x/vio/types.h:250-254); used when timestamp > -1."""
    timestamp: float = -1.0
    x_angle: float = 0.0
    y_angle: float = 0.0


@dataclass
class Measurement:
    """Output of VioUpdater::preProcess (vio_updater.cpp:172-179); tracks are (L,2) arrays, oldest first."""
    timestamp: float = 0.0
    slam_trks: list = field(default_factory=list)
    msckf_trks: list = field(default_factory=list)
    msckf_short_trks: list = field(default_factory=list)
    new_slam_std_trks: list = field(default_factory=list)
    new_msckf_slam_trks: list = field(default_factory=list)
    lost_slam_trk_idxs: list = field(default_factory=list)
    range: RangeMeasurement = None         # VioMeasurement::range (vio/types.h:300)
    sun_angle: SunAngleMeasurement = None  # VioMeasurement::sun_angle (vio/types.h:305)


@dataclass
class PeerState:
    """x::SimpleState (include/x/ekf/simple_state.h:30-75): another agent's snapshot."""
    positions: np.ndarray      # 3*M
    orientations: np.ndarray   # 4*M (x,y,z,w)
    features: np.ndarray       # 3*F
    anchor_idxs: list
    cov: np.ndarray            # N x N
    translation: tuple = (0.0, 0.0, 0.0)


def _csr(tracks):
    off = np.zeros(len(tracks) + 1, dtype=np.int32)
    if tracks:
        off[1:] = np.cumsum([np.asarray(t).shape[0] for t in tracks])
        obs = np.ascontiguousarray(np.vstack([np.asarray(t, dtype=np.float64).reshape(-1, 2) for t in tracks]))
    else:
        obs = np.zeros((0, 2))
    return off, obs


class PackedMeasurement:
    """A Measurement marshalled once into the CSR buffers the C ABI takes (keeps them alive).  pinned=True places the
    observation arrays in page-locked host memory (xb_host_alloc): xb_vio_set_measurement then copies them to the device
    straight from these buffers, without the staging pass."""

    def __init__(self, m: Measurement, pinned=False):
        self.keep = []
        self._pinned = []
        self.c = L.XbMeasurement()
        self.c.timestamp = m.timestamp
        self.h2d_bytes = 0
        lib = L.load() if pinned else None
        for name, trks in (("slam", m.slam_trks), ("msckf", m.msckf_trks), ("msckf_short", m.msckf_short_trks),
                           ("new_slam_std", m.new_slam_std_trks), ("new_msckf_slam", m.new_msckf_slam_trks)):
            off, obs = _csr(trks)
            if pinned and obs.size:
                ptr = lib.xb_host_alloc(obs.nbytes)
                if not ptr:
                    raise MemoryError("xb_host_alloc failed")
                self._pinned.append((lib, ptr))
                buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(obs.size,)).reshape(obs.shape)
                buf[...] = obs
                obs = buf
            self.keep += [off, obs]
            tl = L.XbTrackList(len(trks), L.iptr(off), L.dptr(obs))
            setattr(self.c, name, tl)
            if trks:
                self.h2d_bytes += off.nbytes + obs.nbytes
        lost = np.asarray(m.lost_slam_trk_idxs, dtype=np.int32)
        self.keep.append(lost)
        self.c.n_lost = len(lost)
        self.c.lost_slam_idxs = L.iptr(lost)
        self.range, self.sun_angle = getattr(m, "range", None), getattr(m, "sun_angle", None)

    def __del__(self):
        for lib, ptr in getattr(self, "_pinned", []):
            try:
                lib.xb_host_free(ptr)
            except Exception:
                pass
        self._pinned = []


class PackedMsckfMatches:
    """MsckfMatches (include/x/vision/types.h:83-100) marshalled once into the xb_msckf_match array the C ABI takes -- the
    form in which a C++ front end hands them over anyway; keeps the observation buffers alive."""

    def __init__(self, matches):
        self.keep = []
        self.n = len(matches)
        self.c = Filter._msckf_matches(matches, self.keep)


class Filter:
    """One agent's filter on one GPU: x::Ekf + x::VioUpdater + x::StateManager behind the C ABI."""

    def __init__(self, n_poses_max, n_features_max, **kw):
        lib = L.load()
        self.lib = lib
        cfg = L.XbConfig()
        lib.xb_default_config(C.byref(cfg))
        cfg.n_poses_max, cfg.n_features_max = n_poses_max, n_features_max
        for k, v in kw.items():
            if k == "g":
                for i in range(3):
                    cfg.g[i] = v[i]
            elif not hasattr(cfg, k):
                raise TypeError(f"unknown config field {k}")
            else:
                setattr(cfg, k, v)
        self.cfg = cfg
        self.M, self.F = n_poses_max, n_features_max
        self.N = K_CORE + 6 * self.M + 3 * self.F
        self.LX = xvec_len(self.M, self.F)
        h = C.c_void_p()
        L.check(lib.xb_create(C.byref(cfg), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.xb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- x::Ekf ----------------------------------------------------------------------------
    def initialize_from_state(self, state: State):
        x = L.f64(state.x)
        cov = L.f64(state.cov)
        if x.size != self.LX or cov.shape != (self.N, self.N):
            raise ValueError("init_bfr_mismatch")  # ekf.cpp:50-59
        L.check(self.lib.xb_ekf_initialize_from_state(self.h, L.dptr(x), L.dptr(cov), 0))

    def process_imu(self, t, seq, w_m, a_m, want_state=True):
        out = np.empty(self.LX) if want_state else None
        w, a = L.f64(w_m), L.f64(a_m)
        rc = L.check(self.lib.xb_ekf_process_imu(self.h, float(t), int(seq), L.dptr(w), L.dptr(a), L.dptr(out)))
        if rc == 0:
            return None
        return State(self.M, self.F, out) if want_state else True

    def process_imu_batch(self, samples, want_state=False):
        """A run of Ekf::processImu calls in one (xb_ekf_process_imu_batch): samples = [(t, seq, w_m, a_m), ...].  Returns the
        number of samples that produced a state (and the newest state when want_state)."""
        n = len(samples)
        t = np.ascontiguousarray([s[0] for s in samples], dtype=np.float64)
        sq = np.ascontiguousarray([s[1] for s in samples], dtype=np.uint32)
        w = np.ascontiguousarray([s[2] for s in samples], dtype=np.float64).reshape(-1)
        a = np.ascontiguousarray([s[3] for s in samples], dtype=np.float64).reshape(-1)
        out = np.empty(self.LX) if want_state else None
        rc = L.check(self.lib.xb_ekf_process_imu_batch(self.h, n, L.dptr(t), sq.ctypes.data_as(C.POINTER(C.c_uint)), L.dptr(w),
                                                       L.dptr(a), L.dptr(out)))
        return (rc, State(self.M, self.F, out) if rc > 0 else None) if want_state else rc

    def set_measurement(self, m):
        pm = m if isinstance(m, PackedMeasurement) else PackedMeasurement(m)
        self._pm = pm
        L.check(self.lib.xb_vio_set_measurement(self.h, C.byref(pm.c)))
        rng, sun = pm.range, pm.sun_angle
        if rng is not None or sun is not None:  # VioMeasurement::range / sun_angle (vio/types.h:300-305)
            cr, cs = None, None
            if rng is not None:
                ids = list(rng.tr_feat_ids)
                cr = L.XbRangeMeasurement(rng.timestamp, rng.range, (C.c_double * 2)(*rng.img_pt_n), len(ids),
                                          (C.c_int * 3)(*(ids + [0] * 3)[:3]))
            if sun is not None:
                cs = L.XbSunAngleMeasurement(sun.timestamp, sun.x_angle, sun.y_angle)
            L.check(self.lib.xb_vio_set_sensors(self.h, C.byref(cr) if cr is not None else None,
                                                C.byref(cs) if cs is not None else None))
        return pm

    def process_update_measurement(self, want_state=True):
        out = np.empty(self.LX) if want_state else None
        rc = L.check(self.lib.xb_ekf_process_update(self.h, L.dptr(out)))
        if rc == 0:
            return None
        return State(self.M, self.F, out) if want_state else True

    @staticmethod
    def _peers(peers, keep):
        cp = (L.XbPeerState * max(1, len(peers)))()
        for i, p in enumerate(peers):
            pos, ori, fe, cov = L.f64(p.positions), L.f64(p.orientations), L.f64(p.features), L.f64(p.cov)
            an = np.asarray(p.anchor_idxs, dtype=np.int32)
            keep += [pos, ori, fe, cov, an]
            cp[i].n_poses_max, cp[i].n_features_max = pos.size // 3, fe.size // 3
            cp[i].positions, cp[i].orientations, cp[i].features = L.dptr(pos), L.dptr(ori), L.dptr(fe)
            cp[i].anchor_idxs, cp[i].cov, cp[i].cov_layout = L.iptr(an), L.dptr(cov), 0
            for k in range(3):
                cp[i].translation[k] = p.translation[k]
        return cp

    @staticmethod
    def _msckf_matches(matches, keep):
        """matches: (peer_idx, which, own_track_idx, received_track (L,2)) in list order (vision/types.h:83-100)."""
        cm = (L.XbMsckfMatch * max(1, len(matches)))()
        for j, (pi, which, trk, obs) in enumerate(matches):
            o = np.ascontiguousarray(np.asarray(obs, dtype=np.float64).reshape(-1, 2))
            keep.append(o)
            cm[j].peer, cm[j].which, cm[j].id_current_track, cm[j].n_obs, cm[j].obs = pi, which, trk, o.shape[0], L.dptr(o)
        return cm

    def set_msckf_matches(self, peers, matches):
        """VioUpdater::msckf_matches_ for the next update (vio_updater.cpp:185), peers as SimpleState snapshots."""
        keep = []
        cp = self._peers(peers, keep)
        cm = self._msckf_matches(matches, keep)
        L.check(self.lib.xb_vio_set_msckf_matches(self.h, cp, len(peers), cm, len(matches)))

    def set_msckf_matches_packed(self, gathered_dev_ptr, n_agents, matches):
        """matches: a list of (peer, which, own_track_idx, received_track) or a PackedMsckfMatches (marshalled once)."""
        pm = matches if isinstance(matches, PackedMsckfMatches) else PackedMsckfMatches(matches)
        L.check(self.lib.xb_vio_set_msckf_matches_packed(self.h, C.c_void_p(gathered_dev_ptr), n_agents, pm.c, pm.n))

    def pose_payload_len(self):
        return self.lib.xb_ci_pose_payload_len(self.h)

    def pack_poses(self, dev_ptr, slot=-1):
        """Pack this agent's pose payload (window + 6M x 6M covariance block) into device memory at `dev_ptr`."""
        L.check(self.lib.xb_ci_pack_poses(self.h, slot, C.c_void_p(dev_ptr)))

    def mm_last_gates(self, which=0, n=512):
        out = np.zeros(3 * max(1, n))
        k = L.check(self.lib.xb_mm_last_gates(self.h, which, L.dptr(out), n))
        return out[:3 * k].reshape(-1, 3)

    def apply_ci_lists(self):
        L.check(self.lib.xb_updater_apply_ci_lists(self.h))

    def process_others_measurement(self, t, peers, matches, want_state=True):
        """Ekf::processOthersMeasurement (ekf.cpp:143-176) with SLAM-SLAM matches (peer_idx, current_fid, received_fid)."""
        keep = []
        cp = self._peers(peers, keep)
        cm = (L.XbSlamMatch * max(1, len(matches)))()
        for j, (pi, cur, rcv) in enumerate(matches):
            cm[j].peer, cm[j].current_feature_id, cm[j].received_feature_id = pi, cur, rcv
        out = np.empty(self.LX) if want_state else None
        rc = L.check(self.lib.xb_ekf_process_others(self.h, float(t), cp, len(peers), cm, len(matches), L.dptr(out)))
        if rc == 0:
            return None
        return State(self.M, self.F, out) if want_state else True

    def ci_payload_len(self):
        return self.lib.xb_ci_payload_len(self.h)

    def ci_pack(self, dev_ptr, slot=-1):
        """Pack this agent's compressed CI payload into device memory at `dev_ptr` (e.g. a torch tensor's data_ptr)."""
        L.check(self.lib.xb_ci_pack(self.h, slot, C.c_void_p(dev_ptr)))

    def process_others_packed(self, t, gathered_dev_ptr, n_agents, matches, want_state=True):
        cm = (L.XbSlamMatch * max(1, len(matches)))()
        for j, (pi, cur, rcv) in enumerate(matches):
            cm[j].peer, cm[j].current_feature_id, cm[j].received_feature_id = pi, cur, rcv
        out = np.empty(self.LX) if want_state else None
        rc = L.check(self.lib.xb_ekf_process_others_packed(self.h, float(t), C.c_void_p(gathered_dev_ptr), n_agents, cm,
                                                           len(matches), L.dptr(out)))
        if rc == 0:
            return None
        return State(self.M, self.F, out) if want_state else True

    def ci_last_gates(self, n):
        out = np.zeros(2 * max(1, n))
        k = L.check(self.lib.xb_ci_last_gates(self.h, L.dptr(out), n))
        return out[:2 * k].reshape(-1, 2)

    def get_state(self, slot=-1):
        out = np.empty(self.LX)
        L.check(self.lib.xb_ekf_get_state(self.h, slot, L.dptr(out)))
        return State(self.M, self.F, out)

    def get_covariance(self, slot=-1):
        out = np.empty((self.N, self.N))
        L.check(self.lib.xb_ekf_get_covariance(self.h, slot, L.dptr(out), 0))
        return out

    def newest_slot(self):
        return self.lib.xb_ekf_newest_slot(self.h)

    def synchronize(self):
        L.check(self.lib.xb_synchronize(self.h))

    def set_stream(self, cuda_stream_ptr):
        L.check(self.lib.xb_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    # ---- x::StateManager bookkeeping --------------------------------------------------------
    @property
    def n_poses(self):
        return self.lib.xb_sm_n_poses(self.h)

    @property
    def n_features(self):
        return self.lib.xb_sm_n_features(self.h)

    @property
    def anchor_idxs(self):
        out = np.zeros(max(self.F, 1), dtype=np.int32)
        self.lib.xb_sm_anchor_idxs(self.h, L.iptr(out))
        return out[:self.F].tolist()

    def sm_set(self, n_poses, n_features, anchor_idxs, filled_before):
        a = np.full(max(self.F, 1), -1, dtype=np.int32)
        a[:len(anchor_idxs)] = anchor_idxs
        L.check(self.lib.xb_sm_set(self.h, n_poses, n_features, L.iptr(a), int(filled_before)))

    # ---- stage-level (Updater / VioUpdater / StateManager methods on the work state) -----------
    def work_set(self, state: State):
        x, cov = L.f64(state.x), L.f64(state.cov)
        L.check(self.lib.xb_work_set(self.h, L.dptr(x), L.dptr(cov), 0))

    def work_get(self, with_cov=True):
        x = np.empty(self.LX)
        cov = np.empty((self.N, self.N)) if with_cov else None
        L.check(self.lib.xb_work_get(self.h, L.dptr(x), L.dptr(cov), 0))
        s = State(self.M, self.F, x)
        s.cov = cov
        return s

    def manage(self, lost=()):
        a = np.asarray(list(lost), dtype=np.int32)
        L.check(self.lib.xb_sm_manage(self.h, L.iptr(a), len(a)))

    def construct_update(self, which=0):
        L.check(self.lib.xb_vio_construct_update(self.h, which))

    def reset_correction(self):
        L.check(self.lib.xb_updater_reset_correction(self.h))

    def apply_constructed(self, cov_update=True):
        L.check(self.lib.xb_updater_apply_constructed(self.h, int(cov_update)))

    def apply_update(self, H, res, r_diag, correction_total=None, cov_update=True):
        H, res, r_diag = L.f64(H), L.f64(res), L.f64(r_diag)
        ct = None if correction_total is None else correction_total
        L.check(self.lib.xb_updater_apply_update(self.h, L.dptr(H), L.dptr(res), L.dptr(r_diag), H.shape[0],
                                                 L.dptr(ct), int(cov_update)))

    def apply_ci(self, H, res, S, scaled_block_cols=(), w_result=1.0):
        H, res, S = L.f64(H), L.f64(res), L.f64(S)
        cols = np.asarray(list(scaled_block_cols), dtype=np.int32)
        L.check(self.lib.xb_updater_apply_ci(self.h, L.dptr(H), L.dptr(res), L.dptr(S), H.shape[0], L.iptr(cols),
                                             len(cols), float(w_result)))

    def post_update(self):
        L.check(self.lib.xb_vio_post_update(self.h))

    def updater_update(self):
        L.check(self.lib.xb_updater_update(self.h))

    def update_begin(self, timestamp):
        """First half of Ekf::processUpdateMeasurement (ekf.cpp:183-199): the buffered state closest to `timestamp` becomes
        the work state and is returned (None = std::nullopt)."""
        out = np.empty(self.LX)
        rc = L.check(self.lib.xb_ekf_update_begin(self.h, float(timestamp), L.dptr(out)))
        return State(self.M, self.F, out) if rc else None

    def update_end(self, want_state=True):
        """Second half (ekf.cpp:200-211): write-back and re-propagation; returns the updated state."""
        out = np.empty(self.LX) if want_state else None
        rc = L.check(self.lib.xb_ekf_update_end(self.h, L.dptr(out)))
        if rc == 0:
            return None
        return State(self.M, self.F, out) if want_state else True

    def debug(self, name, count):
        out = np.zeros(int(count))
        n = L.check(self.lib.xb_debug_read(self.h, name.encode(), L.dptr(out), int(count)))
        return out[:n]

    def debug_int(self, name, count):
        out = np.zeros(int(count), dtype=np.int32)
        n = L.check(self.lib.xb_debug_read_int(self.h, name.encode(), L.iptr(out), int(count)))
        return out[:n]

    def profile(self, on=True):
        L.check(self.lib.xb_profile_enable(self.h, int(on)))

    def profile_read(self, reset=True):
        """{stage: (total_ms, count)} accumulated by the per-stage CUDA-event timers."""
        names = (C.c_char_p * 64)()
        ms = np.zeros(64)
        cnt = (C.c_longlong * 64)()
        n = L.check(self.lib.xb_profile_read(self.h, names, L.dptr(ms), cnt, int(reset)))
        return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def kernel_launches(self):
        return int(self.lib.xb_kernel_launches(self.h))
