"""Adapter that gives the fp64 oracle (oracle/) the same call surface as x_multi_agent_b200.Filter,
so that one recorded event stream can be replayed on both.  Test infrastructure only."""
import numpy as np

import oracle
from oracle.updater import VisualMeasurement
from x_multi_agent_b200.filter import State as XState


def to_oracle_state(xs: XState):
    s = oracle.State(xs.M, xs.F)
    s.p, s.v, s.q = xs.p.copy(), xs.v.copy(), xs.q.copy()
    s.b_w, s.b_a = xs.b_w.copy(), xs.b_a.copy()
    s.q_ic, s.p_ic = xs.q_ic.copy(), xs.p_ic.copy()
    s.w_m, s.a_m = xs.w_m.copy(), xs.a_m.copy()
    s.time, s.seq = xs.time, int(xs.x[30])
    s.p_array, s.q_array, s.f_array = xs.p_array.copy(), xs.q_array.copy(), xs.f_array.copy()
    s.cov = np.array(xs.cov, dtype=float).copy()
    return s


def to_oracle_meas(m):
    return VisualMeasurement(timestamp=m.timestamp, slam_trks=list(m.slam_trks), msckf_trks=list(m.msckf_trks),
                             msckf_short_trks=list(m.msckf_short_trks), new_slam_std_trks=list(m.new_slam_std_trks),
                             new_msckf_slam_trks=list(m.new_msckf_slam_trks),
                             lost_slam_trk_idxs=list(m.lost_slam_trk_idxs), **_sensors(m))


def _sensors(m):
    """Copies: the updater marks a sensor measurement as used by resetting its timestamp (vio_updater.cpp:381,402)."""
    from oracle.sensors import RangeMeasurement, SunAngleMeasurement
    out = {}
    r, s = getattr(m, "range", None), getattr(m, "sun_angle", None)
    if r is not None:
        out["range"] = RangeMeasurement(r.timestamp, r.range, tuple(r.img_pt_n), list(r.tr_feat_ids))
    if s is not None:
        out["sun_angle"] = SunAngleMeasurement(s.timestamp, s.x_angle, s.y_angle)
    return out


class OracleFilter:
    def __init__(self, M, F, sigma_img=1.0 / 320.0, rho_0=0.5, sigma_rho_0=0.25, iekf_iter=1, n_slots=250,
                 g=(0.0, 0.0, -9.81), noise=None, sigma_range=0.05):
        self.M, self.F = M, F
        self.upd = oracle.VioUpdaterOracle(M, F, sigma_img, rho_0, sigma_rho_0, iekf_iter, sigma_range=sigma_range)
        self.ekf = oracle.Ekf(self.upd, g, noise or oracle.ImuNoise(), n_slots, oracle.State(M, F))

    def initialize_from_state(self, xs):
        self.upd.sm.clear()
        self.ekf.initialize_from_state(to_oracle_state(xs))

    def process_imu(self, t, seq, w_m, a_m):
        return self.ekf.process_imu(t, seq, w_m, a_m)

    def set_measurement(self, m):
        self.upd.set_measurement(to_oracle_meas(m))

    def process_update_measurement(self):
        return self.ekf.process_update_measurement()

    def newest(self):
        return self.ekf.buf.states[self.ekf.buf.tail]
