"""CPU: host-side logic of the multi-GPU path — index-range sharding and the sum of
per-GPU partials (b200_sum_partials_*, host arithmetic of the product) — against the
oracle, plus a world_size-2 gloo run of the one-process-per-GPU plumbing."""
import os
import socket
import sys

import numpy as np
import pytest

from tests import inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_matches_reference_chunking():
    import legosnark_b200 as lb
    # multiexp.tcc:417-431: one = total/chunks, last chunk takes the remainder
    for n in (0, 1, 7, 8, 9, 1026, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            spans = [lb.shard_range(n, r, world) for r in range(world)]
            covered = [i for lo, hi in spans for i in (lo, hi)]
            assert sum(hi - lo for lo, hi in spans) == n
            if n >= world and world > 1:
                one = n // world
                assert spans[0] == (0, one) and spans[-1] == ((world - 1) * one, n)
            assert covered == sorted(covered)


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_sum_partials_vs_oracle(orc, golden, grp):
    import legosnark_b200 as lb
    g = golden(f"group_{grp}")
    P = g["P"]  # Jacobian, includes zero rows and repeated points
    want = orc.msm(grp, P, inputs.fr_const(P.shape[0], 1), variant=4)
    assert (lb.sum_partials(grp, P) == want).all()
    # doubling and cancellation inside the host adder
    two = np.stack([P[0], g["Q"][0]])  # same point, different Z
    assert (lb.sum_partials(grp, two) == orc.group_op(grp, 3, orc.group_op(grp, 2, P[0:1]))[0]).all()
    canc = np.stack([P[1], g["Q"][1]])  # P + (-P)
    assert (lb.sum_partials(grp, canc) == inputs.zero_point(grp)).all()
    assert (lb.sum_partials(grp, np.zeros((0, P.shape[1]), dtype=np.uint64)) == inputs.zero_point(grp)).all()


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.binding import Checker
    import legosnark_b200 as lb
    from legosnark_b200 import multi
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    orc = Checker("orc")
    P, _ = inputs.bases(orc, "g1", n, seed=71)
    s = inputs.fr_uniform(orc, n, seed=72)
    lo, hi = lb.shard_range(n, rank, world)
    # the per-rank device step is played by the oracle here (CPU test); the sharding,
    # transport and host-side sum are the product's
    partial = orc.msm("g1", P[lo:hi], s[lo:hi])
    total = multi.sharded_multi_exp("g1", partial)
    if rank == 0:
        q.put((total, orc.msm("g1", P, s, chunks=world)))
    dist.barrier()
    dist.destroy_process_group()


def _mailbox_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.binding import Checker
    import legosnark_b200 as lb
    from legosnark_b200 import multi
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    name = f"b200_test_partials_{port}"
    box = multi.HostMailbox(name, rank, world, create=True) if rank == 0 else None
    dist.barrier()
    if box is None:
        box = multi.HostMailbox(name, rank, world, create=False)
    orc = Checker("orc")
    got = []
    for step, grp in enumerate(("g1", "g2", "g1", "g1")):  # several steps: the slots alternate and are reused
        m = n if grp == "g1" else n // 4
        P, _ = inputs.bases(orc, grp, m, seed=81 + step)
        s = inputs.fr_uniform(orc, m, seed=91 + step)
        lo, hi = lb.shard_range(m, rank, world)
        total = multi.sharded_multi_exp(grp, orc.msm(grp, P[lo:hi], s[lo:hi]), mailbox=box)
        got.append((total, orc.msm(grp, P, s, chunks=world)))
    if rank == 0:
        q.put(got)
    dist.barrier()
    box.close(unlink=rank == 0)
    dist.destroy_process_group()


def test_world_size_2_host_mailbox():
    """The partials travel through host shared memory (no collective): what bench.py uses under torchrun."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mailbox_worker, args=(r, 2, port, 203, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for total, want in got:
        assert (total == want).all()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 301, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert (total == want).all()
