"""Multi-GPU plumbing of the ensemble: one process per GPU, members sharded in contiguous blocks, no
data-path collective; one gather of the tallies at the end.

Philox streams are numbered by GLOBAL member index, so results do not depend on the number of ranks.
"""
import numpy as np


def shard_bounds(B, world, rank):
    """Contiguous block [lo, hi) of member indices for `rank`; the remainder is spread over the first ranks."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_tallies(time_local, eo_local, B, group=None):
    """all_gather of time[B_r] (f64) and electrode_occ[B_r,P] (i64) into [B] / [B,P] on every rank.
    Works on CPU tensors (gloo) and CUDA tensors (NCCL).  Uneven shards are padded to the largest block."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    P = eo_local.shape[1]
    sizes = [shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world)]
    mx = max(sizes)
    t_pad = torch.zeros(mx, dtype=torch.float64, device=time_local.device); t_pad[: time_local.shape[0]] = time_local
    e_pad = torch.zeros((mx, P), dtype=torch.int64, device=eo_local.device); e_pad[: eo_local.shape[0]] = eo_local
    ts = [torch.empty_like(t_pad) for _ in range(world)]
    es = [torch.empty_like(e_pad) for _ in range(world)]
    dist.all_gather(ts, t_pad, group=group)
    dist.all_gather(es, e_pad, group=group)
    return torch.cat([t[:n] for t, n in zip(ts, sizes)]), torch.cat([e[:n] for e, n in zip(es, sizes)])


def ensemble_statistics(time_all, eo_all, group_size):
    """Mean / standard error of the currents over the `group_size` seeds of each parameter point."""
    cur = eo_all.astype(np.float64) / time_all[:, None]
    cur = cur.reshape(-1, group_size, cur.shape[1])
    return cur.mean(1), cur.std(1) / np.sqrt(group_size)
