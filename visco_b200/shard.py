"""Multi-GPU plumbing for the hot path: the path shards over (baseline, correlation) matrices with NO exchange
(reference compress_ms.py:588-688 builds one independent task per matrix; decompress_ms.py:188-199 likewise), so each
rank owns a contiguous range of baselines and the only collective is one all-gather of the per-matrix ranks and
statistics at the end of a run (NCCL on GPUs; gloo works for CPU tests)."""
from __future__ import annotations


def shard_baselines(nbl_total: int, world: int, rank: int):
    """Contiguous split of nbl_total baselines over `world` ranks; the first nbl_total % world ranks get one extra.
    Returns (offset, count)."""
    if world < 1 or not (0 <= rank < world) or nbl_total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(nbl_total, world)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def gather_ranks_stats(ranks, stats, counts=None, group=None):
    """All-gather per-matrix `ranks` [B_local] (int32) and `stats` [B_local, 4] (float32) from every rank into global
    arrays ordered by rank. `counts` = matrices per rank (needed when shards are uneven); tensors may live on CPU (gloo)
    or GPU (NCCL)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ranks, stats
    world = dist.get_world_size(group)
    if counts is None:
        counts = [int(ranks.shape[0])] * world
    bmax = max(counts)
    pr = torch.zeros((bmax,), dtype=ranks.dtype, device=ranks.device)
    ps = torch.zeros((bmax, stats.shape[1]), dtype=stats.dtype, device=stats.device)
    pr[: ranks.shape[0]] = ranks
    ps[: stats.shape[0]] = stats
    rl = [torch.empty_like(pr) for _ in range(world)]
    sl = [torch.empty_like(ps) for _ in range(world)]
    dist.all_gather(rl, pr, group=group)
    dist.all_gather(sl, ps, group=group)
    return (torch.cat([r[:c] for r, c in zip(rl, counts)]), torch.cat([s[:c] for s, c in zip(sl, counts)]))
