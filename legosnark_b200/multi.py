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


_bufs = {}  # (limbs, device) -> (pinned in, device in, device out, pinned out): allocated once, the gather runs every step


def gather_partials(partial: np.ndarray, device=None) -> np.ndarray | None:
    """all_gather each rank's partial point; returns (world, limbs) uint64 on every rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(partial, dtype=np.uint64).reshape(1, -1)
    world = dist.get_world_size()
    flat = np.ascontiguousarray(partial, dtype=np.uint64).reshape(-1)
    if dist.get_backend() == "nccl" and device is not None:
        key = (flat.shape[0], str(device))
        if key not in _bufs:
            _bufs[key] = (torch.empty(flat.shape[0], dtype=torch.int64).pin_memory(),
                          torch.empty(flat.shape[0], dtype=torch.int64, device=device),
                          torch.empty((world, flat.shape[0]), dtype=torch.int64, device=device),
                          torch.empty((world, flat.shape[0]), dtype=torch.int64).pin_memory())
        h_in, d_in, d_out, h_out = _bufs[key]
        h_in.numpy()[:] = flat.view(np.int64)
        d_in.copy_(h_in, non_blocking=True)
        dist.all_gather_into_tensor(d_out, d_in)  # one collective of a few hundred bytes (transport, not a data-path collective)
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        return h_out.numpy().view(np.uint64).copy()
    t = torch.from_numpy(flat.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]  # gloo (CPU tests): list form
    dist.all_gather(outs, t)
    return np.stack([o.cpu().numpy().view(np.uint64) for o in outs])


def sharded_multi_exp(group: str, partial: np.ndarray, device=None) -> np.ndarray:
    """Combine this rank's partial with everyone else's: the final, normalised result on every rank."""
    allp = gather_partials(partial, device)
    return sum_partials(group, allp)
