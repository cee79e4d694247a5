"""Inter-agent exchange of the compressed covariance-intersection payload.

One agent per rank / GPU.  Each rank packs its payload on its own GPU (`Filter.ci_pack`), the slots are exchanged
with ONE all-gather (NCCL over NVLink on GPUs; gloo in the CPU tests) and every rank then runs the CI step locally
(`Filter.process_others_packed`).  This is the only collective of the whole path.
"""
import torch
import torch.distributed as dist


def exchange_payloads(local: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather of fixed-size payload slots: returns a [world, len(local)] tensor, slot r = rank r's payload."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = torch.empty((world, local.numel()), dtype=local.dtype, device=local.device)
    if world == 1:
        out[0].copy_(local)
        return out
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


def ring_matches(rank, world, n_features):
    """Synthetic association used by the bench: feature f of this agent <-> feature f of the next agent in the ring."""
    peer = (rank + 1) % world
    return [(peer, f, f) for f in range(n_features)]
