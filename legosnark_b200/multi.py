"""One-process-per-GPU plumbing (torchrun): index-range sharding of one MSM over
the ranks and the host-side sum of the per-rank partial points.

The path has no data-path collective: each rank runs the full
single-GPU pipeline on its slice and returns one normalised point (96 B for G1,
192 B for G2).  Moving those few hundred bytes to rank 0 is the only
communication; it goes through torch.distributed (NCCL on GPUs, gloo in the CPU
tests) purely as transport and the sum itself is b200_sum_partials_* on the host,
mirroring the serial sum of partials at multiexp.tcc:433-438.
"""
from __future__ import annotations

import numpy as np

from . import shard_range, sum_partials  # noqa: F401  (re-exported)


def gather_partials(partial: np.ndarray, device=None) -> np.ndarray | None:
    """all_gather each rank's partial point; returns (world, limbs) uint64 on every rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(partial, dtype=np.uint64).reshape(1, -1)
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    if dist.get_backend() == "nccl":
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t)  # one collective, one device-to-host read
        return out.cpu().numpy().view(np.uint64).reshape(world, -1)
    outs = [torch.empty_like(t) for _ in range(world)]  # gloo (CPU tests): list form
    dist.all_gather(outs, t)
    return np.stack([o.cpu().numpy().view(np.uint64) for o in outs])


def sharded_multi_exp(group: str, partial: np.ndarray, device=None) -> np.ndarray:
    """Combine this rank's partial with everyone else's: the final, normalised result on every rank."""
    allp = gather_partials(partial, device)
    return sum_partials(group, allp)
