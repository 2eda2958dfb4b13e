"""Multi-GPU plumbing.  Environments shard trivially across GPUs (one process
per GPU, contiguous blocks of env ids, Philox keyed on the GLOBAL env id so
results do not depend on the number of shards); there is no collective on the
step path.  The only collective is the optional reduction of the episode
statistics accumulators (K5), once per rollout."""
import math

import torch


def shard_range(total_envs, rank, world_size):
    """Contiguous block [lo, hi) of global env ids owned by ``rank``."""
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_episode_stats(stats, group=None):
    """stats: float64 [8] = episodes, sum length, sum return, min return, max
    return, non-finite, 0, 0 (accumulated atomically by the step kernels).
    All-reduced (SUM / MIN / MAX) when torch.distributed is initialised."""
    s = stats.clone()
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        sums = s[[0, 1, 2, 5]].clone()
        mn, mx = s[3:4].clone(), s[4:5].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        s[[0, 1, 2, 5]] = sums
        s[3], s[4] = mn[0], mx[0]
    v = s.tolist()
    n_ep = v[0]
    return {"episodes": int(n_ep), "mean_length": v[1] / n_ep if n_ep else math.nan,
            "mean_return": v[2] / n_ep if n_ep else math.nan,
            "min_return": v[3] if n_ep else math.nan, "max_return": v[4] if n_ep else math.nan,
            "nonfinite": int(v[5])}
