"""GPU parity tests: the CUDA path (through the C ABI) against the fp64 oracle on identical inputs.

Tolerances (fp64 device arithmetic vs fp64 oracle; the two differ in formulation -- projector/Gram/
Cholesky on the device vs explicit nullspace basis/Householder QR/LU inverse in the oracle -- so the
bar is "agreement to rounding amplified by the problem's conditioning", not bit equality):
  state:       |dp|,|dv| <= 1e-8,  quaternion angle <= 1e-9 rad,  biases/features rel 1e-7
  covariance:  ||dP||_F / ||P||_F <= 1e-8
  gates:       identical inlier masks, gamma rel 1e-7
"""
import numpy as np
import pytest

import oracle
from oracle.updater import apply_ci, apply_update
from oracle_driver import OracleFilter, to_oracle_meas, to_oracle_state
from x_multi_agent_b200 import Filter, State
from x_multi_agent_b200.synth import Scenario, SynthConfig, record, replay

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.linalg.norm(a - b)
    n = max(np.linalg.norm(b), 1e-300)
    return d / n


def quat_angle(q1, q2):
    q1, q2 = q1 / np.linalg.norm(q1), q2 / np.linalg.norm(q2)
    return 2.0 * np.arccos(min(1.0, abs(float(q1 @ q2))))


class Report:
    def __init__(self):
        self.rows, self.bad = [], []

    def check(self, name, val, tol):
        ok = bool(np.isfinite(val)) and val <= tol
        self.rows.append(f"{'ok ' if ok else 'BAD'} {name}: {val:.3e} (tol {tol:.1e})")
        if not ok:
            self.bad.append(name)

    def done(self):
        msg = "\n".join(self.rows)
        print(msg)
        assert not self.bad, "parity failures:\n" + msg


def compare_state(rp, tag, dev: State, ora, M, F, tol_scale=1.0, cov=True):
    rp.check(f"{tag} |dp|", np.linalg.norm(dev.p - ora.p), 1e-8 * tol_scale)
    rp.check(f"{tag} |dv|", np.linalg.norm(dev.v - ora.v), 1e-8 * tol_scale)
    rp.check(f"{tag} angle(q)", quat_angle(dev.q, ora.q), 1e-9 * tol_scale)
    rp.check(f"{tag} b_w", rel(dev.b_w, ora.b_w), 1e-7 * tol_scale)
    rp.check(f"{tag} b_a", rel(dev.b_a, ora.b_a), 1e-7 * tol_scale)
    rp.check(f"{tag} p_array", np.abs(dev.p_array - ora.p_array).max(initial=0.0), 1e-8 * tol_scale)
    rp.check(f"{tag} q_array", np.abs(dev.q_array - ora.q_array).max(initial=0.0), 1e-9 * tol_scale)
    if F:
        rp.check(f"{tag} f_array", np.abs(dev.f_array - ora.f_array).max(initial=0.0), 1e-7 * tol_scale)
    if cov and dev.cov is not None:
        rp.check(f"{tag} cov", rel(dev.cov, ora.cov), 1e-8 * tol_scale)


def make_filter(cfg: SynthConfig, **kw):
    return Filter(cfg.M, cfg.F, max_tracks=max(cfg.K, cfg.n_short, 8), sigma_img=cfg.sigma_img, n_slots=64, **kw)


# ------------------------------------------------------------------------------------------------
def test_propagation_matches_oracle():
    """Propagator::propagateState/propagateCovariance (propagator.cpp:30-205) over a chain of IMU samples."""
    cfg = SynthConfig(M=4, F=3, K=0, seed=3)
    scn = Scenario(cfg)
    s0 = scn.initial_state()
    rng = np.random.default_rng(0)
    n = s0.n_error_states()
    A = rng.normal(size=(n, n)) * 0.05
    s0.cov = A @ A.T + np.diag(rng.uniform(0.01, 0.1, n))
    s0.p_array[:] = rng.normal(size=3 * cfg.M)
    s0.q_array[:] = rng.normal(size=4 * cfg.M)
    s0.f_array[:] = rng.normal(size=3 * cfg.F)
    dev = make_filter(cfg)
    ora = OracleFilter(cfg.M, cfg.F, n_slots=64)
    dev.initialize_from_state(s0)
    ora.initialize_from_state(s0)
    rp = Report()
    for i, (t, seq, w, a) in enumerate([(0.0, 0, *scn.imu_sample(0.0))] + scn.imu_between(0, 3)):
        sd = dev.process_imu(t, seq, w, a)
        so = ora.process_imu(t, seq, w, a)
        assert (sd is None) == (so is None)
        if i in (1, 7, 30):
            sd.cov = dev.get_covariance()
            compare_state(rp, f"imu{i}", sd, so, cfg.M, cfg.F, tol_scale=1e-3)
            rp.check(f"imu{i} P_ii asym (kept)", abs(np.abs(sd.cov[:15, :15] - sd.cov[:15, :15].T).max()
                                                     - np.abs(so.cov[:15, :15] - so.cov[:15, :15].T).max()), 1e-12)
    rp.done()
    dev.close()


def _stage_setup(cfg, n_frames):
    scn = Scenario(cfg)
    ev = record(scn, n_frames)
    last_upd = max(i for i, e in enumerate(ev) if e[0] == "update")
    ora = OracleFilter(cfg.M, cfg.F, sigma_img=cfg.sigma_img, n_slots=64)
    replay(ev[:last_upd], ora)
    m = ev[last_upd][1]
    idx = ora.ekf.buf.closest_idx(m.timestamp)
    assert idx >= 0
    s = ora.ekf.buf.states[idx].copy()
    return ora, m, s


@pytest.mark.parametrize("cfg,frames", [
    (SynthConfig(M=6, F=6, K=12, seed=1), 6),                       # window just full, nothing slid yet
    (SynthConfig(M=6, F=6, K=12, seed=1), 7),                       # first slide + MSCKF-SLAM / std feature init
    (SynthConfig(M=6, F=6, K=12, seed=1, churn=1), 13),             # slide + re-anchoring + feature removal
    (SynthConfig(M=10, F=0, K=50, seed=0), 12),                     # BASELINE cfg-1 shape
    (SynthConfig(M=5, F=4, K=20, seed=5, n_short=3, churn=1), 11),  # short-MSCKF pre-update
    (SynthConfig(M=34, F=4, K=16, seed=3), 37),                     # window > 32 poses: two observations per lane
    (SynthConfig(M=6, F=6, K=0, seed=1, churn=1), 9),               # SLAM-only update (no MSCKF rows at all) with feature loss
])
def test_update_stage_by_stage(cfg, frames):
    """Updater::update (updater.cpp:39-115) split into its stages, each compared with the oracle."""
    ora, m, s = _stage_setup(cfg, frames)
    sm = ora.upd.sm
    dev = make_filter(cfg)
    dev.work_set(State.from_oracle(s))
    dev.sm_set(sm.n_poses, sm.n_features, sm.anchor_idxs, sm.filled_before)
    dev.set_measurement(m)
    rp = Report()
    om = to_oracle_meas(m)
    ora.upd.set_measurement(om)
    corr = np.zeros(s.n_error_states())

    if m.msckf_short_trks:
        h, res, r = ora.upd.construct_short_msckf_update(s)
        sh = ora.upd.last["short"]
        dev.construct_update(1)
        g = dev.debug("gamma0", len(m.msckf_short_trks))
        inl = dev.debug_int("inlier0", len(m.msckf_short_trks))
        rp.check("short gamma", rel(g, sh.gamma), 1e-7)
        rp.check("short inlier mask", float(np.abs(inl - sh.inlier.astype(int)).sum()), 0.0)
        apply_update(s, h, res, r, corr, True)
        dev.apply_constructed(True)
        compare_state(rp, "short", dev.work_get(), s, cfg.M, cfg.F)

    sm.manage(s, list(m.lost_slam_trk_idxs))
    dev.manage(m.lost_slam_trk_idxs)
    compare_state(rp, "manage", dev.work_get(), s, cfg.M, cfg.F, tol_scale=1e-3)
    assert dev.n_poses == sm.n_poses and dev.n_features == sm.n_features
    assert dev.anchor_idxs == list(sm.anchor_idxs)

    h, res, r = ora.upd.construct_update(s)
    dev.construct_update(0)
    ms, mss, sl = ora.upd.last["msckf"], ora.upd.last["msckf_slam"], ora.upd.last["slam"]
    if m.msckf_trks:
        n = len(m.msckf_trks)
        rp.check("msckf ivd", rel(dev.debug("ivd0", 3 * n).reshape(n, 3), ms.features_ivd), 1e-9)
        rp.check("msckf gamma", rel(dev.debug("gamma0", n), ms.gamma), 1e-7)
        rp.check("msckf inlier mask", float(np.abs(dev.debug_int("inlier0", n) - ms.inlier.astype(int)).sum()), 0.0)
    if m.new_msckf_slam_trks:
        n = len(m.new_msckf_slam_trks)
        rp.check("msckf-slam gamma", rel(dev.debug("gamma1", n), mss.gamma), 1e-7)
        rp.check("msckf-slam inlier mask", float(np.abs(dev.debug_int("inlier1", n) - mss.inlier.astype(int)).sum()), 0.0)
    if m.slam_trks:
        n = len(m.slam_trks)
        rp.check("slam gamma", rel(dev.debug("slam_gamma", n), sl.gamma), 1e-7)
        rp.check("slam inlier mask", float(np.abs(dev.debug_int("slam_inlier", n) - sl.inlier.astype(int)).sum()), 0.0)
    # compressed measurement: basis-invariant information  H^T H and H^T r  on the pose columns
    if m.msckf_trks or m.new_msckf_slam_trks:
        M6 = 6 * cfg.M
        rows = np.vstack([ms.jac, mss.jac])
        rr = np.concatenate([ms.res, mss.res])
        G_or = rows[:, 15:15 + M6].T @ rows[:, 15:15 + M6]
        g_or = rows[:, 15:15 + M6].T @ rr
        gp = (M6 + 31) // 32 * 32
        Rg = dev.debug("Rg", gp * gp).reshape(gp, gp)[:M6, :M6]
        Tg = dev.debug("Tg", (gp + 32) * gp).reshape(gp + 32, gp)
        z = Tg[gp, :M6]
        rp.check("gram R^T R", rel(Rg.T @ Rg, G_or), 1e-9)
        rp.check("gram R^T z", rel(Rg.T @ z, g_or), 1e-8)

    corr = np.zeros(s.n_error_states())
    dev.reset_correction()
    if h.size > 0:
        apply_update(s, h, res, r, corr, True)
    dev.apply_constructed(True)
    compare_state(rp, "update", dev.work_get(), s, cfg.M, cfg.F)
    rp.check("correction", rel(dev.debug("corr", len(corr)), corr), 1e-7)

    ora.upd.post_update(s, corr)
    dev.post_update()
    compare_state(rp, "post", dev.work_get(), s, cfg.M, cfg.F)
    assert dev.n_features == sm.n_features and dev.anchor_idxs == list(sm.anchor_idxs)
    P = dev.work_get().cov
    rp.check("P symmetric", np.abs(P - P.T).max() / np.abs(P).max(), 1e-15)
    ev_d, ev_o = np.linalg.eigvalsh(P), np.linalg.eigvalsh(0.5 * (s.cov + s.cov.T))
    rp.check("P spectrum vs oracle", np.abs(ev_d - ev_o).max() / ev_o.max(), 1e-10)
    dev.synchronize()
    rp.done()
    dev.close()


@pytest.mark.parametrize("cfg,frames", [
    (SynthConfig(M=6, F=6, K=12, seed=2, n_short=2, churn=1), 20),
    (SynthConfig(M=10, F=0, K=50, seed=0), 14),
])
def test_full_sequence_through_ekf_api(cfg, frames):
    """Ekf::processImu / processUpdateMeasurement (ekf.cpp:66-255) on a recorded IMU + track stream."""
    scn = Scenario(cfg)
    ev = record(scn, frames)
    ora = OracleFilter(cfg.M, cfg.F, sigma_img=cfg.sigma_img, n_slots=64)
    dev = make_filter(cfg)
    o_states, d_states = [], []
    replay(ev, ora, lambda k, m, st: o_states.append(st.copy()))
    replay(ev, dev, lambda k, m, st: d_states.append(st))
    rp = Report()
    for k in (1, frames // 2, frames - 1):
        compare_state(rp, f"upd{k}", d_states[k], o_states[k], cfg.M, cfg.F, tol_scale=10.0, cov=False)
    dn = dev.get_state()
    dn.cov = dev.get_covariance()
    on = ora.newest()
    compare_state(rp, "newest(repropagated)", dn, on, cfg.M, cfg.F, tol_scale=10.0)
    assert dev.n_poses == ora.upd.sm.n_poses and dev.anchor_idxs == list(ora.upd.sm.anchor_idxs)
    assert dev.kernel_launches() > 0
    dev.synchronize()
    rp.done()
    dev.close()


def test_dense_apply_update_and_ci():
    """Updater::applyUpdate (updater.cpp:117-141) and applyCI (:144-161) with caller-supplied dense matrices."""
    cfg = SynthConfig(M=4, F=5, K=0, seed=7)
    scn = Scenario(cfg)
    s0 = scn.initial_state()
    rng = np.random.default_rng(1)
    n = s0.n_error_states()
    A = rng.normal(size=(n, n)) * 0.03
    s0.cov = A @ A.T + np.diag(rng.uniform(1e-3, 1e-2, n))
    s0.cov[3, 7] += 1e-6  # asymmetric core block as left by Q_d (propagator.cpp:556-570)
    s0.p_array[:] = rng.normal(size=3 * cfg.M)
    q = rng.normal(size=(cfg.M, 4))
    s0.q_array[:] = (q / np.linalg.norm(q, axis=1, keepdims=True)).ravel()
    s0.f_array[:] = rng.normal(size=3 * cfg.F)
    so = to_oracle_state(s0)
    dev = make_filter(cfg)
    dev.work_set(s0)
    rp = Report()
    m = 23
    H = rng.normal(size=(m, n))
    H[:, :15] = 0.0
    res = rng.normal(size=m) * 1e-2
    rd = np.full(m, 1e-4)
    corr_o = rng.normal(size=n) * 1e-3
    corr_d = corr_o.copy()
    apply_update(so, H, res, np.diag(rd), corr_o, True)
    dev.apply_update(H, res, rd, corr_d, True)
    compare_state(rp, "dense", dev.work_get(), so, cfg.M, cfg.F)
    rp.check("dense correction_total", rel(corr_d, corr_o), 1e-9)
    # CI step
    Hc = np.zeros((3, n))
    Hc[:, 15:18] = np.eye(3)
    Hc[:, 15 + 6 * cfg.M:15 + 6 * cfg.M + 3] = rng.normal(size=(3, 3))
    w = 1.25
    cols = [15, 15 + 3 * cfg.M, 15 + 6 * cfg.M]
    Pj = so.cov.copy()
    for c in cols:
        Pj[c:c + 3, c:c + 3] *= w
    S = 1.3 * Hc @ so.cov @ Hc.T + 0.01 * np.eye(3)
    r3 = rng.normal(size=3) * 1e-2
    apply_ci(so, Pj, Hc, r3, S)
    dev.apply_ci(Hc, r3, S, cols, w)
    compare_state(rp, "ci", dev.work_get(), so, cfg.M, cfg.F)
    rp.done()
    dev.close()


def test_cfg2_properties_and_parity():
    """BASELINE cfg-2 (30 poses, 200 SLAM + 800 MSCKF): size-independent properties on the device and one
    full-size update against the oracle started from the device's own steady-state prior."""
    cfg = SynthConfig(M=30, F=200, K=800, seed=0, slam_init_frame=30)
    scn = Scenario(cfg)
    warm = SynthConfig(**{**cfg.__dict__, "K": 40})
    scn_w = Scenario(warm)
    ev = record(scn_w, 33)
    dev = Filter(cfg.M, cfg.F, max_tracks=cfg.K, sigma_img=cfg.sigma_img, n_slots=64)
    replay(ev, dev)
    assert dev.n_poses == cfg.M and dev.n_features == cfg.F
    # one full-size update on both, from the device's prior
    idx_state = dev.get_state()
    scn_w.c.K = cfg.K
    m = scn_w.measurement(33)
    # feed the IMU up to the frame time + latency, as drive() would
    fed = 32 * warm.imu_per_frame + warm.latency_imu
    for i in range(fed + 1, 33 * warm.imu_per_frame + warm.latency_imu + 1):
        t = i * scn_w.dt_imu
        w_m, a_m = scn_w.imu_sample(t)
        dev.process_imu(t, i, w_m, a_m, want_state=False)
    # oracle prior = device state at the slot the update will use
    slot = (dev.newest_slot() - warm.latency_imu) % 64
    prior = dev.get_state(slot)
    assert abs(prior.time - m.timestamp) < 1e-9
    prior.cov = dev.get_covariance(slot)
    so = to_oracle_state(prior)
    upd = oracle.VioUpdaterOracle(cfg.M, cfg.F, cfg.sigma_img)
    upd.sm.n_poses, upd.sm.n_features = dev.n_poses, dev.n_features
    upd.sm.anchor_idxs, upd.sm.filled_before = list(dev.anchor_idxs), True
    upd.set_measurement(to_oracle_meas(m))
    dev.set_measurement(m)
    sd = dev.process_update_measurement()
    sd.cov = dev.get_covariance(slot)
    upd.update(so)
    rp = Report()
    compare_state(rp, "cfg2", sd, so, cfg.M, cfg.F, tol_scale=10.0)
    n = len(m.msckf_trks)
    rp.check("cfg2 msckf inlier mask", float(np.abs(dev.debug_int("inlier0", n) - upd.last["msckf"].inlier.astype(int)).sum()), 0.0)
    rp.check("cfg2 outliers rejected (>=1)", float(upd.last["msckf"].inlier.sum() == n), 0.0)
    P = sd.cov
    rp.check("cfg2 P symmetric", np.abs(P - P.T).max() / np.abs(P).max(), 1e-15)
    ev_d, ev_o = np.linalg.eigvalsh(P), np.linalg.eigvalsh(0.5 * (so.cov + so.cov.T))
    rp.check("cfg2 P spectrum vs oracle", np.abs(ev_d - ev_o).max() / ev_o.max(), 1e-9)
    dev.synchronize()
    rp.done()
    dev.close()


def _two_agents(frames=8):
    """Two agents observing the same SLAM landmarks (same seed -> same landmark ids), different trajectories/noise."""
    out = []
    for a in range(2):
        cfg = SynthConfig(M=6, F=6, K=10, seed=11)          # same seed: same landmark set & feature slot order
        scn = Scenario(cfg)
        scn.phase = scn.phase + 0.3 * a                      # different trajectory per agent
        scn.rng = np.random.Generator(np.random.PCG64(100 + a))
        ev = record(scn, frames)
        ora = OracleFilter(cfg.M, cfg.F, sigma_img=cfg.sigma_img, n_slots=64)
        dev = make_filter(cfg, sigma_landmark=0.3, ci_slam_w=0.1)
        replay(ev, ora)
        replay(ev, dev)
        out.append((cfg, ev, ora, dev))
    return out


def test_slam_slam_covariance_intersection_matches_oracle():
    """Ekf::processOthersMeasurement -> collaborativeUpdate -> MultiSlamUpdate + applyCI
    (ekf.cpp:143-176, updater.cpp:22-36,144-161, multi_slam_update.cpp:61-246, ci.cpp:94-127), fixed weight."""
    from oracle.ci import MultiSlamUpdate, SimpleState, SlamMatch
    from x_multi_agent_b200 import PeerState
    (cfg, ev0, ora0, dev0), (_, ev1, ora1, dev1) = _two_agents()
    sigma_lm, w = 0.3, 0.1
    s1 = ora1.newest()
    peer_o = SimpleState(s1.dynamic_states(), s1.p_array.copy(), s1.q_array.copy(), s1.f_array.copy(), s1.cov.copy(),
                         list(ora1.upd.sm.anchor_idxs))
    peer_d = PeerState(s1.p_array, s1.q_array, s1.f_array, list(ora1.upd.sm.anchor_idxs), s1.cov)
    matches = [(0, f, f) for f in range(cfg.F)] + [(0, 2, 4)]   # last one is a wrong association -> gated out
    t = ora0.newest().time - 0.02                                # a buffered state a few IMU samples back
    # oracle
    def collab(state):
        sm = ora0.upd.sm
        msu = MultiSlamUpdate(sm.camera_attitudes(state), sm.camera_positions(state), state.f_array, sm.anchor_idxs, state.cov,
                              cfg.M, sigma_lm, [SlamMatch(peer_o, c, r) for _, c, r in matches], w)
        collab.msu = msu
        for Pj, H, res, S in zip(msu.P_list, msu.H_list, msu.res_list, msu.S_list):
            apply_ci(state, Pj, H, res, S)
    so = ora0.ekf.process_others_measurement(t, collab)
    sd = dev0.process_others_measurement(t, [peer_d], matches)
    assert so is not None and sd is not None
    rp = Report()
    gates = dev0.ci_last_gates(len(matches))
    rp.check("ci inlier mask", float(np.abs(gates[:, 0] - np.array(collab.msu.inlier, float)).sum()), 0.0)
    rp.check("ci gamma", rel(gates[:, 1], collab.msu.gamma), 1e-8)
    assert sum(collab.msu.inlier) >= 3 and not collab.msu.inlier[-1]
    compare_state(rp, "ci state", sd, so, cfg.M, cfg.F, cov=False)
    dn = dev0.get_state()
    dn.cov = dev0.get_covariance()
    compare_state(rp, "ci newest", dn, ora0.newest(), cfg.M, cfg.F)
    rp.done()
    dev0.close()
    dev1.close()


def test_compressed_ci_payload_exchange_equals_full_state_exchange():
    """exchanging [G_p_f | h P h^T] per feature (13 doubles) gives the same update as shipping the peer's
    SimpleState with its full covariance.  The all-gather is emulated by device copies (one GPU)."""
    import torch
    from x_multi_agent_b200 import PeerState
    (cfg, _, ora0, dev0), (_, _, ora1, dev1) = _two_agents()
    PL = dev0.ci_payload_len()
    gathered = torch.zeros(2, PL, dtype=torch.float64, device="cuda")
    dev0.ci_pack(gathered[0].data_ptr())
    dev1.ci_pack(gathered[1].data_ptr())
    dev0.synchronize()
    dev1.synchronize()
    matches = [(1, f, f) for f in range(cfg.F)]
    t = dev0.get_state().time
    # reference-format exchange on an identical twin of agent 0
    s1 = dev1.get_state()
    peer = PeerState(s1.p_array, s1.q_array, s1.f_array, dev1.anchor_idxs, dev1.get_covariance())
    twin = make_filter(cfg, sigma_landmark=0.3, ci_slam_w=0.1)
    st0 = dev0.get_state()
    st0.cov = dev0.get_covariance()
    for flt in (twin,):
        flt.initialize_from_state(st0)
        flt.sm_set(dev0.n_poses, dev0.n_features, dev0.anchor_idxs, True)
        flt.process_imu(st0.time, 0, st0.w_m, st0.a_m)
    full = twin.process_others_measurement(t, [peer], [(0, c, r) for _, c, r in matches])
    # packed exchange on agent 0 itself (also from its newest state)
    packed = dev0.process_others_packed(t, gathered.data_ptr(), 2, matches)
    rp = Report()
    assert full is not None and packed is not None
    packed.cov = dev0.get_covariance()
    full.cov = twin.get_covariance()
    compare_state(rp, "packed vs full", packed, full, cfg.M, cfg.F)
    g0, g1 = dev0.ci_last_gates(len(matches)), twin.ci_last_gates(len(matches))
    rp.check("gates equal", float(np.abs(g0 - g1).max()), 1e-9)
    assert g0[:, 0].sum() >= 3
    rp.done()
    for d in (dev0, dev1, twin):
        d.close()


def _peer_snapshot(cfg, seed, frames, phase_shift):
    """Another agent (oracle only): its newest state after `frames` updates, as SimpleState (oracle) + PeerState (C ABI)."""
    from oracle.ci import SimpleState
    from x_multi_agent_b200 import PeerState
    scn = Scenario(SynthConfig(M=cfg.M, F=cfg.F, K=cfg.K, seed=seed))
    scn.phase = scn.phase + phase_shift
    ora = OracleFilter(cfg.M, cfg.F, sigma_img=cfg.sigma_img, n_slots=64)
    replay(record(scn, frames), ora)
    s = ora.newest()
    window = list(range(frames - cfg.M, frames))               # camera frames held by the peer's pose window
    so = SimpleState(s.dynamic_states(), s.p_array.copy(), s.q_array.copy(), s.f_array.copy(), s.cov.copy(),
                     list(ora.upd.sm.anchor_idxs))
    sd = PeerState(s.p_array, s.q_array, s.f_array, list(ora.upd.sm.anchor_idxs), s.cov)
    return scn, window, so, sd


@pytest.mark.parametrize("cfg,frames,min_gated,min_inl", [
    (SynthConfig(M=6, F=4, K=14, seed=7, n_short=3), 9, 5, 3),
    (SynthConfig(M=10, F=0, K=50, seed=0), 12, 2, 2),   # cfg-1 shape; the peers' drift turns some matched tracks into outliers
])
def test_multi_uav_msckf_msckf_matches_oracle(cfg, frames, min_gated, min_inl):
    """Updater::update as compiled with -DMULTI_UAV (updater.cpp:39-115): joint triangulation of matched MSCKF tracks,
    stacked 3-row feature blocks, nullspace projection, gate, k-agent covariance intersection, applyCI over the lists,
    then the regular applyUpdate (msckf_update.cpp:65-281,494-501; ci.cpp:49-92; updater.cpp:144-161)."""
    from oracle.ci import MsckfMatch
    scn = Scenario(cfg)
    ev = record(scn, frames)
    last_upd = max(i for i, e in enumerate(ev) if e[0] == "update")
    ora = OracleFilter(cfg.M, cfg.F, sigma_img=cfg.sigma_img, n_slots=64)
    replay(ev[:last_upd], ora)
    m = ev[last_upd][1]
    s = ora.ekf.buf.states[ora.ekf.buf.closest_idx(m.timestamp)].copy()
    sm = ora.upd.sm
    peers = [_peer_snapshot(cfg, 31, frames, 0.25), _peer_snapshot(cfg, 32, frames, -0.2)]
    lms, short_lms = scn.last_msckf_lms, scn.last_short_lms
    rng = np.random.Generator(np.random.PCG64(5))

    def peer_track(p, lm, L):
        pscn, window, _, _ = peers[p]
        return pscn._project(lm, window[len(window) - L:])

    M = cfg.M
    spec = []   # (peer, which, own track, landmark, peer track length)
    spec += [(0, 0, 1, lms[1], M), (1, 0, 1, lms[1], M - 2)]          # two peers on one track (k = 2)
    spec += [(0, 0, 3, lms[3], 3), (1, 0, 4, lms[4], M), (0, 0, 6, lms[6], M - 1)]
    spec += [(1, 0, 8, lms[9], M)]                                     # wrong association: the joint gate rejects it
    if short_lms:
        spec += [(0, 1, 0, short_lms[0], 4), (1, 1, 2, short_lms[2], M)]
    spec += [(0, 0, 10, lms[10], M), (1, 0, 10, lms[10], M)]           # tail of the list: the erase loop visits only the first
    tracks = [peer_track(p, lm, L) for p, _, _, lm, L in spec]
    o_matches = [MsckfMatch(peers[p][2], (which, trk), z) for (p, which, trk, _, _), z in zip(spec, tracks)]
    d_matches = [(p, which, trk, z) for (p, which, trk, _, _), z in zip(spec, tracks)]

    w = 0.1
    dev = make_filter(cfg, multi_uav=1, ci_msckf_w=w)
    dev.work_set(State.from_oracle(s))
    dev.sm_set(sm.n_poses, sm.n_features, sm.anchor_idxs, sm.filled_before)
    dev.set_measurement(m)
    dev.set_msckf_matches([p[3] for p in peers], d_matches)
    dev.updater_update()

    ora.upd.ci_msckf_w = w
    ora.upd.set_measurement(to_oracle_meas(m))
    ora.upd.msckf_matches = list(o_matches)
    ora.upd.update_multi_uav(s)

    rp = Report()
    ms = ora.upd.last["msckf"]
    exp = [(float(j in ms.multi_gate and ms.multi_gate[j][0] < ms.multi_gate[j][1]), ms.multi_gate.get(j, (np.nan, np.nan)))
           for j in range(len(m.msckf_trks)) if ms.n_matched[j] > 0]
    gates = dev.mm_last_gates(0)
    assert len(gates) == len(exp) == 6
    own_ok = [i for i, (_, (g, _)) in enumerate(exp) if np.isfinite(g)]
    rp.check("multi-msckf groups gated", float(len(own_ok) < min_gated), 0.0)
    rp.check("multi-msckf inlier mask", float(np.abs(gates[:, 0] - np.array([e[0] for e in exp])).sum()), 0.0)
    rp.check("multi-msckf gamma", rel(gates[own_ok, 1], [exp[i][1][0] for i in own_ok]), 1e-7)
    rp.check("multi-msckf chi2", rel(gates[own_ok, 2], [exp[i][1][1] for i in own_ok]), 1e-9)
    n_inl = int(sum(e[0] for e in exp))
    assert n_inl >= min_inl and n_inl < len(exp), "scenario must contain accepted and rejected joint updates"
    assert ms.n_matched[10] == 1, "tail-of-list quirk: only the first of the two trailing matches is consumed"
    assert ms.n_matched[1] == 2
    n = len(m.msckf_trks)
    rp.check("msckf gamma (joint triangulation)", rel(dev.debug("gamma0", n), ms.gamma), 1e-7)
    rp.check("msckf inlier mask", float(np.abs(dev.debug_int("inlier0", n) - ms.inlier.astype(int)).sum()), 0.0)
    if m.msckf_short_trks:
        sh = ora.upd.last["short"]
        gs = dev.mm_last_gates(1)
        exps = [sh.multi_gate[j] for j in range(len(m.msckf_short_trks)) if sh.n_matched[j] > 0 and j in sh.multi_gate]
        rp.check("short multi-msckf gamma", rel(gs[np.isfinite(gs[:, 1]), 1], [e[0] for e in exps]), 1e-7)
    compare_state(rp, "multi-uav update", dev.work_get(), s, cfg.M, cfg.F)
    assert dev.n_features == sm.n_features and dev.anchor_idxs == list(sm.anchor_idxs)
    rp.done()
    dev.close()


def test_pose_payload_exchange_equals_full_state_exchange():
    """for MSCKF-MSCKF matches a peer enters only through its pose window and the 6M x 6M pose block of its
    covariance.  Shipping that pose payload (packed on the peer's GPU, xb_ci_pack_poses) gives the same update as
    shipping the whole SimpleState.  Driven through the Ekf-level API (processUpdateMeasurement in a MULTI_UAV build)."""
    import torch
    from x_multi_agent_b200 import PeerState
    cfg = SynthConfig(M=6, F=4, K=12, seed=13)
    frames = 9
    scn0, scn1 = Scenario(cfg), Scenario(SynthConfig(M=6, F=4, K=12, seed=14))
    scn1.phase = scn1.phase + 0.2
    ev0, ev1 = record(scn0, frames), record(scn1, frames - 1)
    last_upd = max(i for i, e in enumerate(ev0) if e[0] == "update")
    kw = dict(multi_uav=1, ci_msckf_w=0.2)
    full, packed, peer = make_filter(cfg, **kw), make_filter(cfg, **kw), make_filter(cfg)
    replay(ev1, peer, want_state=False)
    for flt in (full, packed):
        replay(ev0[:last_upd], flt, want_state=False)
    m = ev0[last_upd][1]
    window1 = list(range(frames - 1 - cfg.M, frames - 1))
    lms = scn0.last_msckf_lms
    matches = [(1, 0, j, scn1._project(lms[j], window1[len(window1) - L:])) for j, L in ((0, 6), (2, 4), (5, 6), (7, 3), (9, 5))]
    PL = peer.pose_payload_len()
    assert PL == 8 + 7 * cfg.M + 36 * cfg.M * cfg.M
    gathered = torch.zeros(2, PL, dtype=torch.float64, device="cuda")
    peer.pack_poses(gathered[1].data_ptr())
    peer.synchronize()
    s1 = peer.get_state()
    ps = PeerState(s1.p_array, s1.q_array, s1.f_array, peer.anchor_idxs, peer.get_covariance())
    pay = gathered[1].cpu().numpy()
    n6 = 6 * cfg.M
    rp = Report()
    rp.check("payload poses", np.abs(pay[8:8 + 3 * cfg.M] - s1.p_array).max(), 0.0)
    rp.check("payload pose covariance block", np.abs(pay[8 + 7 * cfg.M:].reshape(n6, n6) - ps.cov[15:15 + n6, 15:15 + n6]).max(), 0.0)
    full.set_measurement(m)
    full.set_msckf_matches([ps, ps], matches)           # peer index 1, like the gathered slot
    packed.set_measurement(m)
    packed.set_msckf_matches_packed(gathered.data_ptr(), 2, matches)
    a, b = full.process_update_measurement(), packed.process_update_measurement()
    assert a is not None and b is not None
    a.cov, b.cov = full.get_covariance(), packed.get_covariance()
    compare_state(rp, "packed vs full", b, a, cfg.M, cfg.F, tol_scale=1e-3)
    g0, g1 = full.mm_last_gates(0), packed.mm_last_gates(0)
    rp.check("gates equal", float(np.abs(g0 - g1).max()), 1e-9)
    assert len(g0) == len(matches) and g0[:, 0].sum() >= 2
    # matches are consumed by the update they were set for (vio_updater.cpp:185)
    assert full.mm_last_gates(1).size == 0
    rp.done()
    for d in (full, packed, peer):
        d.close()


@pytest.mark.parametrize("cfg,frames", [(SynthConfig(M=6, F=6, K=12, seed=1), 9), (SynthConfig(M=30, F=40, K=60, seed=0), 33)])
def test_tensor_core_downdate_is_fp32_accurate(cfg, frames):
    """Optional tcgen05 (3xTF32, fp32 accumulators in TMEM) covariance downdate vs the default fp64 path.
    The contraction W W^T is fp32-accurate relative to the PRIOR covariance scale (P+ = P - W W^T cancels):
    tolerance ||dP||_F/||P||_F <= 2e-5 and |dP_ij| <= 2e-5 sqrt(Pprior_ii Pprior_jj); measured ~7e-6."""
    ora, m, s = _stage_setup(cfg, frames)
    sm = ora.upd.sm
    devs = [make_filter(cfg, downdate_precision=p) for p in (0, 1)]
    out, prior = [], None
    for dev in devs:
        dev.work_set(State.from_oracle(s))
        dev.sm_set(sm.n_poses, sm.n_features, sm.anchor_idxs, sm.filled_before)
        dev.set_measurement(m)
        dev.manage(m.lost_slam_trk_idxs)
        prior = dev.work_get().cov
        dev.construct_update(0)
        dev.reset_correction()
        dev.apply_constructed(True)
        out.append(dev.work_get())
        dev.synchronize()
    P0, P1 = out[0].cov, out[1].cov
    rp = Report()
    rp.check("tc vs fp64 cov (Frobenius)", rel(P1, P0), 2e-5)
    d = np.sqrt(np.abs(np.diag(prior)))
    act = d > 0
    rp.check("tc vs fp64 cov (component-wise, prior scale)", np.abs((P1 - P0)[np.ix_(act, act)] / np.outer(d[act], d[act])).max(), 2e-5)
    rp.check("tc state identical (correction does not use the downdate)", np.abs(out[0].x - out[1].x).max(), 0.0)
    rp.check("tc P symmetric", np.abs(P1 - P1.T).max(), 0.0)
    rp.done()
    for dev in devs:
        dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,frames", [
    (SynthConfig(M=6, F=6, K=12, seed=2, n_short=2, churn=1), 24),
    (SynthConfig(M=30, F=40, K=60, seed=0, slam_init_frame=30), 40),
])
def test_side_stream_schedule_equals_single_stream(cfg, frames, monkeypatch):
    """The overlapped schedule (SLAM-column half of the update, K-split downdate, build pieces and re-propagation means on
    library-internal side streams) is a re-ordering of independent work: it must give the same states
    and covariance as everything in order on the caller's stream (XB_NO_OVERLAP=1, read at xb_create), update after
    update, also when updates are issued back to back without host synchronisation in between (races would show here)."""
    ev = record(Scenario(cfg), frames)
    results = []
    for no_overlap in ("0", "1"):
        monkeypatch.setenv("XB_NO_OVERLAP", no_overlap)
        dev = make_filter(cfg)
        states = []
        replay(ev, dev, lambda k, m, st: states.append(st))
        P = dev.get_covariance()
        results.append((states, P, dev.get_state()))
        dev.synchronize()
        dev.close()
    monkeypatch.delenv("XB_NO_OVERLAP")
    (s_a, P_a, n_a), (s_b, P_b, n_b) = results
    rp = Report()
    assert len(s_a) == len(s_b) and len(s_a) > 0
    worst = max(np.abs(a.x - b.x).max() for a, b in zip(s_a, s_b))
    rp.check("states, all updates (abs)", worst, 1e-11)
    rp.check("newest state", np.abs(n_a.x - n_b.x).max(), 1e-11)
    rp.check("newest covariance", rel(P_a, P_b), 1e-11)
    rp.done()


@pytest.mark.gpu
def test_constructed_update_survives_calls_between_construct_and_apply():
    """The C ABI allows xb_updater_reset_correction / xb_sm_manage-style calls between constructUpdate and applyUpdate (the
    reference's Updater::update does its own sequencing, updater.cpp:39-115).  The early side-stream part of a constructed
    update must then be discarded and rebuilt (invalidate_early): same result as the plain sequence."""
    cfg = SynthConfig(M=6, F=6, K=12, seed=1, churn=1)
    ora, m, s = _stage_setup(cfg, 13)
    out = []
    for variant in range(2):
        dev = make_filter(cfg)
        dev.work_set(State.from_oracle(s))
        sm = ora.upd.sm
        dev.sm_set(sm.n_poses, sm.n_features, sm.anchor_idxs, sm.filled_before)
        dev.set_measurement(m)
        dev.manage(m.lost_slam_trk_idxs)
        dev.reset_correction()
        dev.construct_update(0)
        dev.apply_constructed(False)          # IEKF-style first pass: correction_total becomes non-zero, covariance untouched
        dev.construct_update(0)               # second pass: the early part is built with that correction_total ...
        if variant == 1:
            dev.reset_correction()            # ... which the caller now resets before applying
        else:
            dev.reset_correction()
            dev.construct_update(0)           # reference sequence: reset first, then construct
        dev.apply_constructed(True)
        out.append(dev.work_get())
        dev.synchronize()
        dev.close()
    rp = Report()
    rp.check("state", np.abs(out[0].x - out[1].x).max(), 1e-12)
    rp.check("covariance", rel(out[0].cov, out[1].cov), 1e-12)
    rp.done()


@pytest.mark.gpu
def test_pinned_measurement_buffers_take_the_direct_copy_path():
    """xb_vio_set_measurement copies observation arrays in page-locked memory (xb_host_alloc) without staging; results
    are identical to the staged path."""
    from x_multi_agent_b200 import PackedMeasurement
    cfg = SynthConfig(M=10, F=8, K=200, seed=4)
    ev = record(Scenario(cfg), 14)
    out = []
    for pinned in (False, True):
        dev = make_filter(cfg)
        keep = []
        for e in ev:
            if e[0] == "init":
                dev.initialize_from_state(e[1])
            elif e[0] == "imu":
                dev.process_imu(*e[1:], want_state=False)
            else:
                pm = PackedMeasurement(e[1], pinned=pinned)
                keep.append(pm)           # pinned buffers must outlive the update that reads them
                dev.set_measurement(pm)
                dev.process_update_measurement(want_state=False)
        st = dev.get_state()
        st.cov = dev.get_covariance()
        out.append(st)
        dev.synchronize()
        dev.close()
    assert np.array_equal(out[0].x, out[1].x) and np.array_equal(out[0].cov, out[1].cov)


@pytest.mark.gpu
def test_consecutive_empty_updates_match_the_oracle():
    """An update without measurement rows runs StateManager::manage but no applyUpdate (updater.cpp:106), so the
    reference's covariance is not symmetrised: the asymmetry of its Q_d (propagator.cpp:207-840) spreads from the core
    block into every clone added meanwhile (state_manager.cpp:273-349) and P_vi != P_iv^T is propagated
    (propagator.cpp:197-203).  The device carries that state exactly (second strip per ring slot, column-side products
    in manage, the general applyUpdate of k_general.cu for the first update with rows).  Six consecutive empty updates
    (window filling without any track), then feature initialisation and SLAM-only updates: regular tolerances."""
    cfg = SynthConfig(M=6, F=6, K=0, seed=1, churn=1)
    ev = record(Scenario(cfg), 14)
    ora = OracleFilter(cfg.M, cfg.F, sigma_img=cfg.sigma_img, n_slots=64)
    dev = make_filter(cfg)
    o_states, d_states = [], []
    replay(ev, ora, lambda k, m, st: o_states.append(st.copy()))
    replay(ev, dev, lambda k, m, st: d_states.append(st))
    rp = Report()
    for k in range(len(d_states)):
        compare_state(rp, f"upd{k}", d_states[k], o_states[k], cfg.M, cfg.F, tol_scale=10.0, cov=False)
    dn = dev.get_state()
    dn.cov = dev.get_covariance()
    compare_state(rp, "newest", dn, ora.newest(), cfg.M, cfg.F, tol_scale=10.0)
    dev.synchronize()
    rp.done()
    dev.close()


@pytest.mark.gpu
def test_late_peer_message_is_applied_to_an_older_buffered_state():
    """Ekf::processOthersMeasurement with a timestamp several visual updates back (ekf.cpp:143-176): the reference
    applies any timestamp inside its state buffer and re-propagates everything after it.  The device keeps a buffered
    state usable for n_generations - 1 later covariance updates (default 31) and answers XB_E_STALE -- not a silent
    nullopt -- beyond that."""
    from oracle.ci import MultiSlamUpdate, SimpleState, SlamMatch
    from x_multi_agent_b200 import PeerState
    from x_multi_agent_b200.lib import XbError
    (cfg, ev0, ora0, dev0), (_, ev1, ora1, dev1) = _two_agents(frames=12)
    sigma_lm, w = 0.3, 0.1
    s1 = ora1.newest()
    peer_o = SimpleState(s1.dynamic_states(), s1.p_array.copy(), s1.q_array.copy(), s1.f_array.copy(), s1.cov.copy(),
                         list(ora1.upd.sm.anchor_idxs))
    peer_d = PeerState(s1.p_array, s1.q_array, s1.f_array, list(ora1.upd.sm.anchor_idxs), s1.cov)
    matches = [(0, f, f) for f in range(cfg.F)]
    t = ora0.newest().time - 0.17          # 34 IMU states / three visual updates back

    def collab(state):
        sm = ora0.upd.sm
        msu = MultiSlamUpdate(sm.camera_attitudes(state), sm.camera_positions(state), state.f_array, sm.anchor_idxs, state.cov,
                              cfg.M, sigma_lm, [SlamMatch(peer_o, c, r) for _, c, r in matches], w)
        for Pj, H, res, S in zip(msu.P_list, msu.H_list, msu.res_list, msu.S_list):
            apply_ci(state, Pj, H, res, S)
    so = ora0.ekf.process_others_measurement(t, collab)
    sd = dev0.process_others_measurement(t, [peer_d], matches)
    assert so is not None and sd is not None and abs(sd.time - so.time) < 1e-12 and so.time < ora0.newest().time - 0.15
    rp = Report()
    compare_state(rp, "late ci state", sd, so, cfg.M, cfg.F, cov=False)
    dn = dev0.get_state()
    dn.cov = dev0.get_covariance()
    compare_state(rp, "late ci newest", dn, ora0.newest(), cfg.M, cfg.F)
    rp.done()
    dev0.close()
    dev1.close()
    # a filter with only three generations: the same late message now addresses a recycled generation
    cfg3 = SynthConfig(M=6, F=6, K=10, seed=11)
    dev = make_filter(cfg3, sigma_landmark=0.3, ci_slam_w=0.1, n_generations=3)
    replay(ev0, dev)
    with pytest.raises(XbError) as ei:
        dev.process_others_measurement(t, [peer_d], matches)
    assert ei.value.code == -7
    dev.close()


@pytest.mark.gpu
def test_ci_fusion_at_cfg2_size_with_four_peers():
    """BASELINE cfg-3 shape: SLAM-SLAM covariance-intersection fusion at cfg-2 dimensions (30-pose window, 200 SLAM
    features, N = 795) with four peers and 16 matches per peer, plus one wrong association per peer.
    Peers and prior come from device filters that ran the fill sequence; the fusion itself (MultiSlamUpdate, pair
    fuseCI, applyCI in list order, "last match wins" covariance: multi_slam_update.cpp:61-246, ci.cpp:94-127,
    updater.cpp:22-36,144-161) is compared with the oracle on the same inputs."""
    from oracle.ci import MultiSlamUpdate, SimpleState, SlamMatch
    from x_multi_agent_b200 import PeerState
    M, F, n_peers, per_peer = 30, 200, 4, 16
    sigma_lm, w = 0.1, 0.1
    agents = []
    for a in range(n_peers + 1):
        cfg = SynthConfig(M=M, F=F, K=12, seed=40 + a, slam_init_frame=M, slam_lm_seed=4242, slam_msckf_init_frac=1.0)
        dev = Filter(M, F, max_tracks=16, sigma_img=cfg.sigma_img, n_slots=64, sigma_landmark=sigma_lm, ci_slam_w=w,
                     oc_projection=0)
        replay(record(Scenario(cfg), 36), dev)
        assert dev.n_features == F
        agents.append((cfg, dev))
    cfg, own = agents[0]
    peers_d, peers_o = [], []
    for _, dev in agents[1:]:
        s = dev.get_state()
        P = dev.get_covariance()
        peers_d.append(PeerState(s.p_array.copy(), s.q_array.copy(), s.f_array.copy(), dev.anchor_idxs, P))
        peers_o.append(SimpleState(s.x[0:16].copy(), s.p_array.copy(), s.q_array.copy(), s.f_array.copy(), P.copy(),
                                   list(dev.anchor_idxs)))
    # every peer matches a different block of 16 features; the last entry of each block is a wrong association
    matches = []
    for p in range(n_peers):
        for q in range(per_peer):
            f = p * per_peer + q
            matches.append((p, f, f if q < per_peer - 1 else (f + 97) % F))
    slot = own.newest_slot()
    prior = own.get_state(slot)
    prior.cov = own.get_covariance(slot)
    so = to_oracle_state(prior)
    sm_anchor = list(own.anchor_idxs)
    quats = [so.q_array[4 * i:4 * i + 4] for i in range(own.n_poses)]
    poss = [so.p_array[3 * i:3 * i + 3] for i in range(own.n_poses)]
    msu = MultiSlamUpdate(quats, poss, so.f_array, sm_anchor, so.cov, M, sigma_lm,
                          [SlamMatch(peers_o[p], c, r) for p, c, r in matches], w)
    for Pj, H, res, S in zip(msu.P_list, msu.H_list, msu.res_list, msu.S_list):
        apply_ci(so, Pj, H, res, S)
    sd = own.process_others_measurement(prior.time, peers_d, matches)
    assert sd is not None
    rp = Report()
    gates = own.ci_last_gates(len(matches))
    rp.check("ci inlier mask", float(np.abs(gates[:, 0] - np.array(msu.inlier, float)).sum()), 0.0)
    rp.check("ci gamma", rel(gates[:, 1], msu.gamma), 1e-8)
    n_inl = int(sum(msu.inlier))
    print(f"cfg-2-size CI: {n_inl}/{len(matches)} matches fused")
    assert n_inl >= len(matches) // 2, "the fusion arithmetic must actually run on accepted matches"
    assert not any(msu.inlier[p * per_peer + per_peer - 1] for p in range(n_peers)), "wrong associations must be gated out"
    compare_state(rp, "ci state", sd, so, M, F, cov=False)
    sd.cov = own.get_covariance(slot)
    rp.check("ci covariance", rel(sd.cov, so.cov), 1e-8)
    rp.done()
    for _, dev in agents:
        dev.close()


@pytest.mark.gpu
def test_two_filters_on_one_gpu_run_concurrently():
    """Two agents' filters on ONE GPU, driven from two host threads at the same time (ctypes releases the GIL): their
    tile-Cholesky dataflow launches are cooperative (all CTAs of a launch resident or none), so neither can starve the
    other.  Both must reproduce their single-filter results bit for bit."""
    import threading
    cfgs = [SynthConfig(M=30, F=40, K=60, seed=21 + a, slam_init_frame=30) for a in range(2)]
    evs = [record(Scenario(c), 40) for c in cfgs]

    def run(c, ev, out):
        dev = Filter(c.M, c.F, max_tracks=c.K, sigma_img=c.sigma_img, n_slots=64)
        xs = []
        replay(ev, dev, lambda k, m, st: xs.append(st.x.copy()))
        dev.synchronize()
        out.append(np.vstack(xs))
        dev.close()

    alone = [[], []]
    for a in range(2):
        run(cfgs[a], evs[a], alone[a])
    both = [[], []]
    th = [threading.Thread(target=run, args=(cfgs[a], evs[a], both[a])) for a in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
        assert not t.is_alive(), "a filter hung while the other one was running"
    for a in range(2):
        assert np.array_equal(alone[a][0], both[a][0])


# ---- range (laser range finder) and sun-sensor rows ---------------------------------------------------
SENSOR_CASES = {
    # rows <= N + 1: no QR compression in the reference, the sensor rows keep their own variances
    "slam_only_no_qr": (dict(M=5, F=6, K=0, seed=5, churn=1, range_every=1, sun_every=3), 16, 1.0),
    # rows > N + 1: the reference QR-compresses and weights EVERY row sigma_img^2 (vio_updater.cpp:490-508); the range row
    # then acts like a 3 mm measurement and the later updates are ill-conditioned (round-off x10 per update in the
    # reference binary and the oracle alike, tests/test_ref_pinning.py): loose final tolerance, tight early one
    "msckf_slam_qr": (dict(M=6, F=6, K=14, seed=11, n_short=2, churn=1, range_every=1, sun_every=2), 16, 1e3),
    # sun sensor before the first SLAM feature exists (sparse part = the two sun rows only) and IEKF: the sensor rows
    # enter the first iteration only (vio_updater.cpp:381, 402)
    "sun_first_iekf2": (dict(M=8, F=8, K=10, seed=5, n_short=3, churn=2, slam_init_frame=3, range_every=2, sun_every=1), 14, 10.0),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(SENSOR_CASES))
@pytest.mark.parametrize("no_overlap", ["0", "1"])
def test_range_and_sun_rows_match_oracle(case, no_overlap, monkeypatch):
    """RangeUpdate (range_update.cpp:61-265) + SolarUpdate (solar_update.cpp:39-94) stacked under the visual rows
    (vio_updater.cpp:352-403): gates, every update's state and the final covariance against the oracle (which is pinned
    to the reference's own sources for exactly these cases), on the side-stream schedule and in order."""
    kw, frames, loose = SENSOR_CASES[case]
    cfg = SynthConfig(**kw)
    iekf = 2 if case == "sun_first_iekf2" else 1
    ev = record(Scenario(cfg), frames)
    ora = OracleFilter(cfg.M, cfg.F, sigma_img=cfg.sigma_img, n_slots=64, iekf_iter=iekf, sigma_range=cfg.sigma_range)
    monkeypatch.setenv("XB_NO_OVERLAP", no_overlap)
    dev = make_filter(cfg, iekf_iter=iekf, sigma_range=cfg.sigma_range)
    o_states, d_states, o_gate, d_gate = [], [], [], []

    def on_oracle(k, m, st):
        o_states.append(st.copy())
        ru = ora.upd.last.get("range")
        o_gate.append((k, ru.inlier, ru.gamma) if ru is not None else None)

    def on_dev(k, m, st):
        d_states.append(st)
        used = m.range is not None and len(m.slam_trks) > 0
        d_gate.append((k, bool(dev.debug_int("range_inlier", 1)[0]), dev.debug("range_gamma", 1)[0]) if used else None)

    replay(ev, ora, on_oracle)
    replay(ev, dev, on_dev)
    rp = Report()
    n_gated = 0
    for go, gd in zip(o_gate, d_gate):
        assert (go is None) == (gd is None)
        if go is not None:
            n_gated += 1
            assert go[1] == gd[1], f"range gate differs at update {go[0]}"
            rp.check(f"range gamma upd{go[0]}", abs(gd[2] - go[2]) / abs(go[2]), 1e-7 * loose)
    assert n_gated >= 3 and any(g[1] for g in o_gate if g is not None)
    first = next(i for i, g in enumerate(o_gate) if g is not None and g[1])
    compare_state(rp, f"upd{first}(first range inlier)", d_states[first], o_states[first], cfg.M, cfg.F, tol_scale=10.0, cov=False)
    compare_state(rp, f"upd{frames - 1}", d_states[-1], o_states[-1], cfg.M, cfg.F, tol_scale=10.0 * loose, cov=False)
    dn = dev.get_state()
    dn.cov = dev.get_covariance()
    compare_state(rp, "newest(repropagated)", dn, ora.newest(), cfg.M, cfg.F, tol_scale=10.0 * loose)
    dev.synchronize()
    rp.done()
    dev.close()


@pytest.mark.gpu
def test_imu_batch_equals_per_sample_calls():
    """xb_ekf_process_imu_batch == the same samples through Ekf::processImu one by one (ekf.cpp:66-140): identical ring
    bookkeeping (a repeated timestamp is skipped, an accelerometer spike repeats the previous reading), estimates equal to
    round-off (the batch runs the re-propagation kernels: prefix products instead of a step-by-step recurrence), and a
    whole sequence with updates in between ends in the same state and covariance."""
    cfg = SynthConfig(M=6, F=6, K=12, seed=4, n_short=2, churn=1)
    ev = record(Scenario(cfg), 12)
    a_dev, b_dev = make_filter(cfg), make_filter(cfg)
    sa, sb = [], []
    pending = []

    def flush():
        if pending:
            # a duplicate timestamp and an accelerometer spike inside the batch
            t0, q0, w0, a0 = pending[len(pending) // 2]
            batch = list(pending) + [(t0, q0, w0, a0)]
            batch.sort(key=lambda s: s[0])
            spike = len(batch) - 1
            batch[spike] = (batch[spike][0], batch[spike][1], batch[spike][2], np.array([500.0, 0.0, 0.0]))
            for (t, q, w, a) in batch:
                a_dev.process_imu(t, q, w, a, want_state=False)
            n = b_dev.process_imu_batch(batch)
            assert n == len(pending), "the duplicate timestamp must be skipped"
            pending.clear()

    for e in ev:
        if e[0] == "init":
            a_dev.initialize_from_state(e[1]); b_dev.initialize_from_state(e[1])
        elif e[0] == "imu":
            pending.append((e[1], e[2], e[3], e[4]))
        else:
            flush()
            assert a_dev.newest_slot() == b_dev.newest_slot()
            rp0 = np.abs(a_dev.get_state().x - b_dev.get_state().x).max()
            assert rp0 < 1e-12, rp0
            for d, out in ((a_dev, sa), (b_dev, sb)):
                d.set_measurement(e[1])
                out.append(d.process_update_measurement())
    rp = Report()
    rp.check("states, all updates (abs)", max(np.abs(x.x - y.x).max() for x, y in zip(sa, sb)), 1e-11)
    rp.check("newest covariance", rel(a_dev.get_covariance(), b_dev.get_covariance()), 1e-11)
    rp.done()
    a_dev.close(); b_dev.close()
