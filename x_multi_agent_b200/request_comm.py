"""Request/response exchange of the covariance-intersection payload (the reference's -DREQUEST_COMM build).

reference: src/x/vio/vio.cpp:455-496 (getDescriptors / processOtherRequests), src/x/place_recognition/vlad.cpp:68-75
(VLAD score), src/x/place_recognition/database.cpp:30-61 (candidate search, 15-keyframe database),
src/x/vio/vio_updater.cpp:451-484 (keyframe selection).  In the reference an agent broadcasts a 81 x 32-byte VLAD
descriptor of its current view; a peer answers with a stored keyframe (state snapshot + tracks) only if the descriptor
scores above a threshold against one of its keyframes, and at most once per requesting agent and keyframe.  That
policy is what cuts the inter-agent traffic (README.md:98-99).

Here the descriptor matching itself (ORB + DBoW3 vocabulary) is out of scope; this module is the policy and the
transport: the request is the same 2592-byte bit vector, the answer is the agent's compressed CI payload
(`Filter.ci_pack`, 8 + 13 F doubles instead of the reference's full SimpleState), requests travel in ONE small
all-gather and answers as point-to-point NCCL send/recv between exactly the pairs whose request was accepted.
"""
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

VLAD_CLUSTERS, VLAD_BYTES = 81, 32          # k^L = 3^4 words x 32-byte ORB descriptors (vlad.cpp:24-31)
VLAD_LEN = VLAD_CLUSTERS * VLAD_BYTES       # 2592 bytes on the wire
_POPCOUNT = np.array([bin(i).count("1") for i in range(256)], dtype=np.int64)


def vlad_score(x: np.ndarray, y: np.ndarray) -> float:
    """VLAD::computeScore (vlad.cpp:68-75): (bits - hamming(x xor y)) / bits."""
    bits = 8 * VLAD_LEN
    ham = int(_POPCOUNT[np.bitwise_xor(x.reshape(-1), y.reshape(-1))].sum())
    return (bits - ham) / bits


@dataclass
class Keyframe:
    vlad: np.ndarray                       # uint8 [VLAD_LEN]
    payload: torch.Tensor                  # the CI payload snapshot taken when the keyframe was selected
    time: float = 0.0
    sent_to: set = field(default_factory=set)   # Keyframe::findOtherUavId / setOtherUavId


class KeyframeDatabase:
    """Database (database.cpp): at most 15 keyframes, oldest dropped first."""

    def __init__(self, pr_score_thr: float, max_keyframes: int = 15):
        self.thr, self.max_keyframes, self.keyframes = pr_score_thr, max_keyframes, []

    def add(self, kf: Keyframe):               # Database::addKeyframe (database.cpp:52-61)
        self.keyframes.append(kf)
        if len(self.keyframes) > self.max_keyframes:
            self.keyframes.pop(0)

    def find_candidate(self, uav_id: int, query: np.ndarray):
        """Database::findCandidate (database.cpp:30-50): best-scoring keyframe above the threshold that has not been
        sent to `uav_id` yet; it is marked as sent."""
        best, score = None, 0.0
        for kf in self.keyframes:
            if uav_id in kf.sent_to:
                continue
            s = vlad_score(query, kf.vlad)
            if s > self.thr and s > score:
                best, score = kf, s
        if best is not None:
            best.sent_to.add(uav_id)
        return best, score


class KeyframeSelector:
    """Keyframe selection of VioUpdater::postUpdate (vio_updater.cpp:451-484): after more than 10 frames, when the
    agent has moved by more than 15 % of the mean SLAM feature depth and more than 10 tracks are alive."""

    def __init__(self):
        self.frames_min_distance = 0
        self.last_pose = np.zeros(3)

    def step(self, position, inverse_depths, n_tracks):
        take = False
        if self.frames_min_distance > 10:
            rho = np.asarray(inverse_depths, dtype=float).reshape(-1, 3)[:, 2]
            # the reference sums every third entry starting at index 3 and divides by the feature count (:459-464)
            r = rho[1:]
            med_depth = float(np.abs(1.0 / r[r > 0.001]).sum() / max(len(rho), 1)) if len(rho) else 0.0
            diff = np.linalg.norm(np.asarray(position) - self.last_pose)
            if med_depth > 0.0 and diff / med_depth > 0.15 and n_tracks > 10:
                take = True
                self.last_pose = np.array(position, dtype=float)
                self.frames_min_distance = 0
        self.frames_min_distance += 1
        return take


def exchange_request_response(request: torch.Tensor, db: KeyframeDatabase, payload_len: int, group=None):
    """One request/response round.  `request`: this agent's VLAD descriptor, uint8 [VLAD_LEN] on the payload device.
    Every rank (1) all-gathers the requests, (2) looks for a keyframe for every requester (processOtherRequests,
    vio.cpp:462-496), (3) all-gathers the world x world accept matrix so that both ends of a pair agree, and
    (4) exchanges the answers point to point.  Returns ({peer: payload tensor}, stats)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = request.device
    reqs = torch.empty((world, VLAD_LEN), dtype=torch.uint8, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(reqs.view(-1), request.contiguous().view(-1), group=group)
    else:
        reqs[0].copy_(request)
    reqs_h = reqs.cpu().numpy()
    answers = {}
    accept = torch.zeros(world, dtype=torch.uint8, device=dev)
    for peer in range(world):
        if peer == rank:
            continue
        kf, _ = db.find_candidate(peer, reqs_h[peer])
        if kf is not None:
            answers[peer] = kf.payload
            accept[peer] = 1
    acc = torch.empty((world, world), dtype=torch.uint8, device=dev)    # acc[a][b] = 1: a answers b's request
    if world > 1:
        dist.all_gather_into_tensor(acc.view(-1), accept, group=group)
    else:
        acc[0].copy_(accept)
    acc_h = acc.cpu().numpy()
    received, ops = {}, []
    for peer in range(world):
        if peer == rank:
            continue
        if acc_h[rank][peer]:
            ops.append(dist.P2POp(dist.isend, answers[peer].contiguous(), peer, group))
        if acc_h[peer][rank]:
            buf = torch.empty(payload_len, dtype=torch.float64, device=dev)
            received[peer] = buf
            ops.append(dist.P2POp(dist.irecv, buf, peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    stats = {"requests_bytes": VLAD_LEN * (world - 1), "answers_sent": int(acc_h[rank].sum()),
             "answers_received": int(acc_h[:, rank].sum()), "answer_bytes": payload_len * 8}
    return received, stats
