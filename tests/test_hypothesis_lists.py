"""Host logic of the per-env hypothesis LISTS (batched_agents._alive_rows: three and four agents) on CPU tensors:
the rows that survive pruning, per env, against the dense mask the two-agent path keeps (bd:200-256, 792-886)."""
import types

import numpy as np
import pytest
import torch

from gym_cooking_b200 import batched_agents
from gym_cooking_b200.delegation_planner import hypothesis_space


def _dense(T, ok):
    okp = torch.cat([ok, torch.ones((ok.shape[0], 1), dtype=torch.bool)], dim=1)
    alive = T.static_ok[None, :].clone().expand(ok.shape[0], -1).clone()
    for e in range(T.E):
        alive &= okp[:, T.ent_lidx[:, e]]
    return alive


@pytest.mark.parametrize("H,E,L,N,p_ok", [(300, 3, 40, 500, 0.8), (5000, 4, 90, 200, 0.9), (17, 2, 5, 64, 0.5), (64, 4, 130, 50, 0.97)])
def test_alive_rows_equal_the_dense_mask(H, E, L, N, p_ok):
    rng = np.random.RandomState(H + L)
    ent = rng.randint(L + 1, size=(H, E))  # L = "no entry" (always fine)
    T = types.SimpleNamespace(H=H, E=E, ent_lidx=torch.from_numpy(ent), static_ok=torch.from_numpy(rng.rand(H) < 0.9))
    ok = torch.from_numpy(rng.rand(N, L) < p_ok)
    ok[: N // 4] = ok[0]  # many envs share a bit pattern, as in a real batch
    ok[-1] = False
    rows, count = batched_agents.BatchedDelegation._alive_rows(None, T, ok)
    alive = _dense(T, ok)
    assert torch.equal(count, alive.sum(1))
    assert rows.shape[1] == max(int(count.max()), 1)
    for n in range(N):
        want = alive[n].nonzero()[:, 0]
        assert torch.equal(rows[n, : want.numel()], want), n          # front-packed, ascending
        assert bool((rows[n, want.numel():] == H).all()), n           # padded with the row nobody has


def test_alive_rows_on_a_real_three_agent_table():
    """the bd table of three agents over four subtasks: with only subtask 0 doable for everybody, what survives is
    every allocation that hands subtask 0 to one agent or one pair and leaves the others idle"""
    names = ["agent-1", "agent-2", "agent-3"]
    subtasks = ["A", "B", "C", "D"]
    allocs = list(dict.fromkeys(tuple(a) for a in hypothesis_space("bd", names[0], names, subtasks)))
    agsets = [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2)]
    lid = {(s, ag): k for k, (s, ag) in enumerate((s, ag) for s in range(4) for ag in agsets)}
    L, E = len(lid), max(len(a) for a in allocs)
    ent = np.full((len(allocs), E), L, dtype=np.int64)
    static_ok = np.ones(len(allocs), dtype=bool)
    for h, alloc in enumerate(allocs):
        for e, t in enumerate(alloc):
            ag = tuple(sorted(int(nm.split("-")[1]) - 1 for nm in t.subtask_agent_names))
            if t.subtask is None:
                static_ok[h] &= len(ag) == 1
            else:
                ent[h, e] = lid[(subtasks.index(t.subtask), ag)]
        if all(t.subtask is None for t in alloc):
            static_ok[h] = False
    T = types.SimpleNamespace(H=len(allocs), E=E, ent_lidx=torch.from_numpy(ent), static_ok=torch.from_numpy(static_ok))
    ok = torch.zeros((2, L), dtype=torch.bool)
    ok[0, [lid[(0, ag)] for ag in agsets]] = True  # env 0: subtask A doable by every agent set; env 1: nothing is
    rows, count = batched_agents.BatchedDelegation._alive_rows(None, T, ok)
    assert int(count[1]) == 0 and bool((rows[1] == T.H).all())
    survivors = [allocs[int(h)] for h in rows[0, : int(count[0])]]
    assert len(survivors) == 6  # A to one of three agents or one of three pairs, nobody else works
    for alloc in survivors:
        working = [t for t in alloc if t.subtask is not None]
        assert len(working) == 1 and working[0].subtask == "A"
