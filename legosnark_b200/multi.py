"""One-process-per-GPU plumbing (torchrun): index-range sharding of one MSM over
the ranks and the host-side sum of the per-rank partial points.

The path has no data-path collective: each rank runs the full
single-GPU pipeline on its slice and returns one normalised point (96 B for G1,
192 B for G2).  Moving those few hundred bytes between the ranks of one box is
the only communication.  `HostMailbox` does it through a page of host shared
memory (/dev/shm): no NCCL, no device round trip — the north star's "partials
are summed on the host".  `gather_partials` is the torch.distributed transport
kept for launches without a shared /dev/shm (gloo in the CPU tests).  Either way
the sum itself is b200_sum_partials_* on the host, mirroring the serial sum of
partials at multiexp.tcc:433-438.
"""
from __future__ import annotations

import mmap
import os
import time

import numpy as np

from . import shard_range, sum_partials  # noqa: F401  (re-exported)


class HostMailbox:
    """Exchange of per-rank partial points through host shared memory.

    One file in /dev/shm holds, per rank, a sequence number and two payload slots (even / odd steps) of
    `limbs` uint64 each.  exchange() publishes this rank's partial for the step, waits until every rank has
    published the same step and returns all partials (world, limbs).  A rank can only reach step s + 2 (which
    reuses the slot of step s) after every rank has entered step s + 1, i.e. finished reading step s.
    Rank 0 creates the file before the launcher's barrier; the others open it after."""
    SLOT = 64  # uint64 per rank: [0] seq, [8..8+limbs) even payload, [36..36+limbs) odd payload

    def __init__(self, name: str, rank: int, world: int, create: bool, limbs: int = 24):
        assert limbs <= 24
        self.rank, self.world, self.limbs, self.seq = rank, world, limbs, 0
        self.path = os.path.join("/dev/shm", name)
        size = world * self.SLOT * 8
        if create:
            with open(self.path, "wb") as f:
                f.write(b"\0" * size)
        self.f = open(self.path, "r+b")
        self.mm = mmap.mmap(self.f.fileno(), size)
        self.a = np.frombuffer(self.mm, dtype=np.uint64).reshape(world, self.SLOT)

    def exchange(self, partial: np.ndarray, timeout_s: float = 120.0) -> np.ndarray:
        self.seq += 1
        off = 8 if self.seq % 2 == 0 else 36
        flat = np.ascontiguousarray(partial, dtype=np.uint64).reshape(-1)
        m = flat.shape[0]  # 12 (G1) or 24 (G2) limbs; every rank sends the same group in a given step
        assert m <= self.limbs
        self.a[self.rank, off:off + m] = flat
        self.a[self.rank, 0] = self.seq  # published after the payload (x86 keeps the store order)
        t0 = time.perf_counter()
        while int(self.a[:, 0].min()) < self.seq:
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError("HostMailbox: a rank did not publish its partial")
        return self.a[:, off:off + m].copy()

    def close(self, unlink: bool = False):
        self.a = None
        try:
            self.mm.close()
            self.f.close()
        except BufferError:
            pass
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


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


def sharded_multi_exp(group: str, partial: np.ndarray, device=None, mailbox: HostMailbox | None = None) -> np.ndarray:
    """Combine this rank's partial with everyone else's: the final, normalised result on every rank."""
    allp = mailbox.exchange(partial) if mailbox is not None else gather_partials(partial, device)
    return sum_partials(group, allp)
