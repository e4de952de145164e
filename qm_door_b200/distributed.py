"""Multi-GPU plumbing of the batch axis (BASELINE config 4 / north_star): independent MPC problems are sharded
contiguously across ranks (no data-path collective), and the per-rank policy shard is all-gathered once per cycle
(torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests). The reference has no multi-process numerics
(SURVEY.md §2.3); this axis is what the B200 build adds."""
import torch
import torch.distributed as dist

POLICY_WIDTH = 61      # per node: interpolation time, x*[30], u*[30]  (feed-forward policy, task.info:90 useFeedbackPolicy false)


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` problems for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_policy(t, x, u, out=None):
    """(t[B,N], x[B,N,30], u[B,N,30]) -> [B,N,61] contiguous buffer for the collective."""
    B, N = t.shape
    if out is None:
        out = torch.empty(B, N, POLICY_WIDTH, dtype=t.dtype, device=t.device)
    out[..., 0] = t
    out[..., 1:31] = x
    out[..., 31:61] = u
    return out


def unpack_policy(buf):
    return buf[..., 0], buf[..., 1:31], buf[..., 31:61]


def init_comm(ctx):
    """Bootstrap of the context's own NCCL communicator over an initialised torch.distributed group: rank 0 draws the unique id
    (qmb200_nccl_unique_id), the group broadcasts its 128 bytes, every rank calls qmb200_comm_init. The collective itself
    (qmb200_allgather_policy) then runs inside the library, on the context's communication stream."""
    import qm_door_b200 as q
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    if world == 1:
        ctx.enable_policy_buffer()
        return
    box = [q.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)


def allgather_ragged(shard, total):
    """Shards from shard_range differ by one problem when total % world != 0, and the tensor collective needs equal shapes:
    every rank pads its [b,N,61] shard to ceil(total / world) problems, the padding is cut out after the gather.
    Returns [total,N,61] in batch order."""
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    per = -(-total // world)
    padded = torch.zeros((per,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
    padded[:shard.shape[0]] = shard
    g = allgather_policy(padded)
    parts = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        parts.append(g[r, :hi - lo])
    return torch.cat(parts, dim=0)


def allgather_policy(shard, gathered=None):
    """All ranks contribute an equally shaped [B,N,61] shard; returns [world,B,N,61] (torch.distributed: the host-logic path of
    the CPU tests; on GPUs the library's own qmb200_allgather_policy is the product path)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if gathered is None:
        gathered = torch.empty((world,) + tuple(shard.shape), dtype=shard.dtype, device=shard.device)
    if world == 1:
        gathered[0].copy_(shard)
    else:
        # concatenated-along-dim-0 output form: accepted by both NCCL and gloo
        dist.all_gather_into_tensor(gathered.view((-1,) + tuple(shard.shape[1:])), shard.contiguous())
    return gathered
