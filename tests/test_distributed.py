"""N>1 host logic on CPU: world_size 2, gloo. Each rank solves its contiguous shard of the batch (here with the CPU port
standing in for the device — the sharding / packing / all-gather plumbing is what is under test) and the gathered policy
must equal the single-process solution of the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import abi_fill
    from qm_door_b200 import distributed as D, workload
    W = workload.Workload(total, horizon=0.1, dt=0.01, seed=77)
    lo, hi = D.shard_range(total, rank, world)
    cp = abi_fill.CPort(W.model, W.problem, W.solver, hi - lo)
    out = cp.cycle(np.zeros(hi - lo), W.x0[lo:hi], W.events[lo:hi], W.modes[lo:hi], W.nevents[lo:hi], W.target_t[lo:hi], W.target_x[lo:hi])
    shard = D.pack_policy(torch.from_numpy(out["t"]), torch.from_numpy(out["x"]), torch.from_numpy(out["u"]))
    gathered = D.allgather_ragged(shard, total)        # shard sizes differ by one when total % world != 0
    if rank == 0:
        q.put(gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    from qm_door_b200.distributed import shard_range
    for total in (1, 7, 8, 1024, 5632):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(5632, 3, 8) == (2112, 2816)          # config 4: 704 problems per rank


import pytest


@pytest.mark.parametrize("total", [6, 7])
def test_two_rank_gloo_allgather_equals_single_process(descs, total):
    from oracle import abi_fill
    from qm_door_b200 import distributed as D, workload
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    W = workload.Workload(total, horizon=0.1, dt=0.01, seed=77)
    cp = abi_fill.CPort(W.model, W.problem, W.solver, total)
    ref = cp.cycle(np.zeros(total), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
    t, x, u = D.unpack_policy(torch.from_numpy(gathered).reshape(total, -1, D.POLICY_WIDTH))
    assert np.array_equal(t.numpy(), ref["t"]) and np.array_equal(x.numpy(), ref["x"]) and np.array_equal(u.numpy(), ref["u"])
