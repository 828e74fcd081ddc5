"""Env-batch sharding across the GPUs of one box.

Episodes are independent, so rank r simply owns a contiguous range of GLOBAL env indices; the
philox action stream is keyed on the global index, which makes every result independent of the
number of GPUs.  The only collective on this path is the final sum of the episode-statistics
vector (gc_stats_reduce -> all_reduce): NCCL on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist

STATS_FIELDS = ("episodes", "successes", "sum_t_done", "sum_collisions", "running")


def shard_range(n_total, rank, world_size):
    """[lo, hi) of the envs rank `rank` owns; sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_stats(stats):
    """Sum a per-rank statistics vector (int64[GC_STATS_LEN]) over all ranks, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def stats_dict(stats):
    v = stats.tolist() if isinstance(stats, torch.Tensor) else list(stats)
    out = dict(zip(STATS_FIELDS, v[:5]))
    out["t_histogram"] = v[5:133]
    out["completed_subtasks"] = v[133] if len(v) > 133 else 0
    return out


MIXED_TOTALS = ("envs", "agent_steps", "posterior_updates", "delivered", "planning_states_solved", "planner_lookups",
                "completed_subtasks")


def reduce_mixed_totals(totals, device=None):
    """Whole-job totals of batched_agents.run_mixed when every rank ran its own shard of episodes (cfg-5): the counts
    are summed over ranks, the wall time is the slowest rank's.  Returns a dict with MIXED_TOTALS + "seconds"."""
    counts = torch.tensor([int(totals[k]) for k in MIXED_TOTALS], dtype=torch.int64, device=device)
    seconds = torch.tensor([float(totals["seconds"])], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        dist.all_reduce(seconds, op=dist.ReduceOp.MAX)
    out = dict(zip(MIXED_TOTALS, counts.cpu().tolist()))
    out["seconds"] = float(seconds.item())
    return out
