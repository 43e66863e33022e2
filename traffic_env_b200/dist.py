"""Env-parallel sharding across GPUs: one process per GPU, no per-step collective.

Env instances never interact (all coupling is inside one env, roadgraph.py:36-39), so rank g owns
the contiguous block of global env ids [g*E/G, (g+1)*E/G).  The Philox arrival stream and the reset
phases are keyed by the GLOBAL env id, so trajectories do not depend on the number of ranks.  The
only collective is a sum of a few counters (episode returns as defined by util.py:68-94, episode
and overflow counts, vehicle updates) per reporting interval: NCCL over NVLink on GPUs, gloo in the
CPU tests.
"""
import torch
import torch.distributed as dist

STAT_KEYS = ("ticks", "actor_steps", "vehicle_updates", "overflows", "cars_generated", "episodes",
             "return_sum", "disc_return_sum", "seq_fallback_ticks", "cars_exited")


def shard_range(total_envs, rank, world):
    """[begin, end) of the global env ids owned by `rank`; remainders go to the low ranks."""
    base, rem = divmod(int(total_envs), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_stats(stats, device=None, group=None):
    """Sum a VecTrafficEnv.stats() dict over all ranks (all_reduce of one float64 vector)."""
    if not (dist.is_available() and dist.is_initialized()):
        return dict(stats)
    t = torch.tensor([float(stats[k]) for k in STAT_KEYS], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = {}
    for k, v in zip(STAT_KEYS, t.tolist()):
        out[k] = v if k in ("return_sum", "disc_return_sum") else int(round(v))
    return out


def mean_episode_return(stats, discounted=True):
    """The quantity the reference prints per episode (util.py:68-94), averaged over closed episodes."""
    n = stats["episodes"]
    if n == 0:
        return float("nan")
    return (stats["disc_return_sum"] if discounted else stats["return_sum"]) / n
