"""Multi-process path on CPU (gloo, world size 2): sharded env ranges + the one collective of
the path (sum of episode statistics) reproduce the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
import gym_cooking_b200 as gcb
from gym_cooking_b200 import sharding

N_TOTAL, N_AGENTS, LEVEL, STEPS, MAX_T = 6001, 2, "open-divider_tomato", 30, 20


def _stats(lo, hi):
    """episode statistics of envs [lo, hi) after STEPS philox steps, from the CPU oracle"""
    lv = O.parse_level(gcb.levels.level_text(LEVEL), MAX_T)
    st = O.reset_state(lv, N_AGENTS, hi - lo)
    rd, coll, _ = O.rollout_batch(lv, st, N_AGENTS, STEPS, env0=lo, seed=99)
    t = (st[:, 0] >> 24) & 127
    done = st[:, 0] >> 31
    out = np.zeros(134, dtype=np.int64)
    out[0], out[1], out[2] = hi - lo, int(((done == 1) & (t < MAX_T)).sum()), int(t[done == 1].sum())
    out[3], out[4] = int(coll.sum()), int((done == 0).sum())
    out[5:133] = np.bincount(t[done == 1], minlength=128)
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(N_TOTAL, rank, world)
    stats = torch.from_numpy(_stats(lo, hi))
    sharding.reduce_stats(stats)
    if rank == 0:
        q.put(stats.tolist())
    dist.barrier()
    dist.destroy_process_group()


def _totals_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = dict(envs=100 + rank, agent_steps=1000 * (rank + 1), posterior_updates=7 * rank, delivered=90 - rank,
                planning_states_solved=50 + rank, planner_lookups=400, completed_subtasks=3 * (rank + 2),
                seconds=1.5 + 2.0 * rank, per_level=[])
    got = sharding.reduce_mixed_totals(mine)
    if rank == 0:
        q.put(got)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduction_of_the_cfg5_totals():
    """what bench.py reports for cfg-5 at N > 1: counts summed over ranks, wall time of the slowest rank"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_totals_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == dict(envs=201, agent_steps=3000, posterior_updates=7, delivered=179, planning_states_solved=101,
                       planner_lookups=800, completed_subtasks=15, seconds=3.5)
    single = sharding.reduce_mixed_totals(dict(envs=5, agent_steps=6, posterior_updates=7, delivered=1,
                                               planning_states_solved=2, planner_lookups=3, completed_subtasks=4, seconds=0.25))
    assert single["agent_steps"] == 6 and single["seconds"] == 0.25  # no process group: unchanged


def test_shard_ranges_cover_exactly():
    for n, w in ((10, 3), (8, 8), (5, 8), (1 << 23, 8), (6001, 2)):
        r = [sharding.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_gloo_reduction_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == _stats(0, N_TOTAL).tolist()
    d = sharding.stats_dict(got)
    assert d["episodes"] == N_TOTAL and d["running"] + sum(d["t_histogram"]) == N_TOTAL
