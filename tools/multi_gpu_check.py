"""Multi-GPU check of the library's collective (run under gpurun --gpus 2/4/8 with torchrun):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py
Every rank solves its contiguous shard of one batch (config-2 shape at a short horizon), qmb200_allgather_policy gathers the
packed policy over NCCL on the context's communication stream, and every rank checks the gathered policy against the
single-process solution of the whole batch computed on its own GPU (bit-identical: problems are independent)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qm_door_b200 as q  # noqa: E402
from qm_door_b200 import distributed as D, workload  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per = 24
    total = per * world
    W = workload.Workload(total, horizon=0.3, dt=0.01, seed=99)
    lo, hi = D.shard_range(total, rank, world)
    ctx = q.MpcContext(W.model, W.problem, W.solver, per, device=local)
    D.init_comm(ctx)
    full = q.MpcContext(W.model, W.problem, W.solver, total, device=local)
    dev = torch.device("cuda", local)
    gathered = [torch.zeros(world, per, W.solver.max_nodes, 61, dtype=torch.float64, device=dev) for _ in range(2)]
    sl = slice(lo, hi)
    worst = 0
    for c in range(4):
        t0 = np.full(total, 0.01 * c)
        ctx.cycle(t0[sl], W.x0[sl], W.events[sl], W.modes[sl], W.nevents[sl], W.target_t[sl], W.target_x[sl])
        ctx.allgather_policy(gathered[c & 1])                     # runs beside the reference solve below
        ref = full.cycle(t0, W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
        ctx.comm_sync()
        g = gathered[c & 1].reshape(total, W.solver.max_nodes, 61).cpu().numpy()
        for b in range(total):
            n = ref["n"][b]
            ok = (np.array_equal(g[b, :n, 0], ref["t"][b, :n]) and np.array_equal(g[b, :n, 1:31], ref["x"][b, :n])
                  and np.array_equal(g[b, :n, 31:], ref["u"][b, :n]))
            worst += int(not ok)
    bad = torch.tensor([worst], device=dev)
    dist.all_reduce(bad)
    if rank == 0:
        print("multi_gpu_check: world %d, %d problems, 4 cycles, mismatching problems: %d" % (world, total, int(bad.item())))
    ctx.close(); full.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(bad.item()) else 0)


if __name__ == "__main__":
    main()
