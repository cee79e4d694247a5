"""N > 1 path on CPU (gloo, world_size 2): the payload exchange plumbing, and that the compressed CI payload gives the
same fusion as shipping the peer's full SimpleState (oracle arithmetic on both sides)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _agent(rank, frames=8):
    from oracle_driver import OracleFilter
    from x_multi_agent_b200.synth import Scenario, SynthConfig, record, replay
    cfg = SynthConfig(M=6, F=6, K=10, seed=11)
    scn = Scenario(cfg)
    scn.phase = scn.phase + 0.3 * rank
    scn.rng = np.random.Generator(np.random.PCG64(100 + rank))
    ora = OracleFilter(cfg.M, cfg.F, n_slots=64)
    ev = record(scn, frames)
    replay(ev, ora)
    return cfg, ora, scn, ev[-1][1]


def _peer_truth(rank):
    """Analytic trajectory of another agent (no filter run): used to project this agent's landmarks into its camera."""
    from x_multi_agent_b200.synth import Scenario, SynthConfig
    scn = Scenario(SynthConfig(M=6, F=6, K=10, seed=11))
    scn.phase = scn.phase + 0.3 * rank
    scn.rng = np.random.Generator(np.random.PCG64(200 + rank))
    return scn


def _worker(rank, world, port, q):
    sys.path.insert(0, os.fspath(ROOT))
    sys.path.insert(0, os.fspath(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.ci import MultiSlamUpdate, SimpleState, SlamMatch, multi_slam_from_payload, pack_payload
    from oracle.updater import apply_ci
    from x_multi_agent_b200.ci import exchange_payloads, ring_matches
    cfg, ora, scn, last_meas = _agent(rank)
    s, sm = ora.newest(), ora.upd.sm
    local = torch.from_numpy(pack_payload(s, sm, cfg.F))
    gathered = exchange_payloads(local).numpy()
    assert gathered.shape == (world, local.numel())
    assert np.array_equal(gathered[rank], local.numpy())
    matches = ring_matches(rank, world, cfg.F)
    quats, poss = sm.camera_attitudes(s), sm.camera_positions(s)
    comp = multi_slam_from_payload(quats, poss, s.f_array, sm.anchor_idxs, s.cov, cfg.M, 0.3, gathered, matches, 0.1)
    # full-state exchange: ship the whole SimpleState (object all-gather), as the reference does
    full_states = [None] * world
    dist.all_gather_object(full_states, (s.dynamic_states(), s.p_array, s.q_array, s.f_array, s.cov, list(sm.anchor_idxs)))
    peers = [SimpleState(*fs) for fs in full_states]
    msu = MultiSlamUpdate(quats, poss, s.f_array, sm.anchor_idxs, s.cov, cfg.M, 0.3,
                          [SlamMatch(peers[p], c, r) for p, c, r in matches], 0.1)
    a, b = s.copy(), s.copy()
    for Pj, H, res, S in zip(msu.P_list, msu.H_list, msu.res_list, msu.S_list):
        apply_ci(a, Pj, H, res, S)
    for Pj, H, res, S in zip(comp["P"], comp["H"], comp["res"], comp["S"]):
        apply_ci(b, Pj, H, res, S)
    ok = (list(msu.inlier) == list(comp["inlier"]) and np.allclose(msu.gamma, comp["gamma"], rtol=1e-10)
          and np.linalg.norm(a.cov - b.cov) <= 1e-12 * np.linalg.norm(a.cov) and np.linalg.norm(a.p - b.p) < 1e-12
          and sum(msu.inlier) >= 2)
    # MSCKF-MSCKF matches: the pose payload (window + 6M x 6M covariance block) replaces the full SimpleState
    from oracle.ci import MsckfMatch, MultiMsckfUpdate, pack_pose_payload, peer_from_pose_payload
    pose_local = torch.from_numpy(pack_pose_payload(s, cfg.M))
    pose_gathered = exchange_payloads(pose_local).numpy()
    peer = (rank + 1) % world
    truth = _peer_truth(peer)
    window = list(range(8 - cfg.M, 8))
    trks, lms = last_meas.msckf_trks, scn.last_msckf_lms
    ptracks = {j: truth._project(lms[j], window[cfg.M - L:]) for j, L in ((0, 6), (3, 4), (7, 5))}
    lists = []
    for peer_state in (peers[peer], peer_from_pose_payload(pose_gathered[peer], cfg.F)):
        mm = [MsckfMatch(peer_state, j, z) for j, z in ptracks.items()]
        mu = MultiMsckfUpdate(trks, list(range(len(trks))), quats, poss, s.cov, cfg.M, 1.0 / 320.0, mm, 0.1)
        lists.append(mu)
    a_, b_ = lists
    ok_mm = (len(a_.S_list) == len(b_.S_list) and len(a_.multi_gate) == 3
             and all(np.array_equal(x, y) for x, y in zip(a_.S_list, b_.S_list))
             and all(np.array_equal(x, y) for x, y in zip(a_.H_list, b_.H_list))
             and np.array_equal(a_.jac, b_.jac))
    ok = ok and ok_mm
    payload_bytes, full_bytes = local.numel() * 8, sum(np.asarray(x).nbytes for x in full_states[rank][:5])
    q.put((rank, bool(ok), payload_bytes, full_bytes))
    dist.destroy_process_group()


def test_two_rank_gloo_payload_exchange_and_ci_equivalence():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    assert all(pb * 10 < fb for _, _, pb, fb in res)   # the compressed slot is >10x smaller than the SimpleState


# ---- request / response exchange (the reference's -DREQUEST_COMM build, vio.cpp:455-496) ------------------------
def _rc_worker(rank, world, port, q):
    sys.path.insert(0, os.fspath(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from x_multi_agent_b200.request_comm import (VLAD_LEN, Keyframe, KeyframeDatabase, exchange_request_response, vlad_score)
    rng = np.random.Generator(np.random.PCG64(5))
    scene_a = rng.integers(0, 256, VLAD_LEN, dtype=np.uint8)
    scene_b = rng.integers(0, 256, VLAD_LEN, dtype=np.uint8)      # an unrelated place: ~50 % of the bits differ

    def noisy(v, seed):
        r = np.random.Generator(np.random.PCG64(seed))
        return np.bitwise_xor(v, np.packbits(r.uniform(size=8 * VLAD_LEN) < 0.03))

    PL = 21
    db = KeyframeDatabase(pr_score_thr=0.9)
    # rank 0 has seen place A, rank 1 has seen place B and place A
    db.add(Keyframe(noisy(scene_a, 10 + rank), torch.full((PL,), 100.0 + rank, dtype=torch.float64)))
    if rank == 1:
        db.add(Keyframe(noisy(scene_b, 20), torch.full((PL,), 200.0, dtype=torch.float64)))
    ok = abs(vlad_score(scene_a, scene_a) - 1.0) < 1e-15 and 0.4 < vlad_score(scene_a, scene_b) < 0.6
    ok = ok and vlad_score(scene_a, noisy(scene_a, 1)) > 0.95
    # round 1: rank 0 asks about place A, rank 1 asks about place B (which rank 0 never saw)
    req = torch.from_numpy(noisy(scene_a if rank == 0 else scene_b, 30 + rank))
    got, st = exchange_request_response(req, db, PL)
    if rank == 0:
        ok = ok and list(got) == [1] and float(got[1][0]) == 101.0 and st["answers_sent"] == 0
    else:
        ok = ok and got == {} and st["answers_sent"] == 1
    # round 2: the same request again -- the keyframe was already sent to that agent (database.cpp:33-36)
    got2, st2 = exchange_request_response(req, db, PL)
    ok = ok and got2 == {} and st2["answers_sent"] == 0 and st2["answers_received"] == 0
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo_request_response_policy():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_keyframe_selector_follows_the_reference_rule():
    """vio_updater.cpp:451-484: no keyframe during the first 10 frames, then one whenever the agent moved by more than
    15 % of the mean feature depth with more than 10 live tracks."""
    from x_multi_agent_b200.request_comm import KeyframeSelector
    sel = KeyframeSelector()
    ivd = np.tile([0.0, 0.0, 0.1], 20)          # 20 features at 10 m
    taken = [sel.step(np.array([0.2 * k, 0.0, 0.0]), ivd, n_tracks=50) for k in range(40)]
    assert not any(taken[:11]) and taken[11]
    nxt = taken.index(True, 12)
    assert nxt - 11 == 11                       # the frame counter restarts at every keyframe
    assert not any(sel.step(np.array([100.0 + k, 0, 0]), ivd, n_tracks=5) for k in range(30))   # not "worthy"
