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


def allgather_policy(shard, gathered=None):
    """All ranks contribute an equally shaped [B,N,61] shard; returns [world,B,N,61]."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if gathered is None:
        gathered = torch.empty((world,) + tuple(shard.shape), dtype=shard.dtype, device=shard.device)
    if world == 1:
        gathered[0].copy_(shard)
    else:
        # concatenated-along-dim-0 output form: accepted by both NCCL and gloo
        dist.all_gather_into_tensor(gathered.view((-1,) + tuple(shard.shape[1:])), shard.contiguous())
    return gathered
