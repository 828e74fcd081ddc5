"""Host logic of the batched delegation loop on CPU tensors: the static hypothesis / likelihood-row
tables (batched_agents._ObserverTables) against `hypothesis_space`, and the alive-mask / prior
arithmetic against a plain Python restatement of prune_subtask_allocs + get_spatial_priors
(bd:200-256, 296-369).  No kernel runs here."""
import itertools
import types

import numpy as np
import torch

from gym_cooking_b200 import batched_agents as ba
from gym_cooking_b200 import recipe_planner as rp
from gym_cooking_b200.delegation_planner import hypothesis_space


def _owner(n_agents=2, recipe=("Salad",), kinds=("Tomato", "Lettuce", "Plate", "Plate")):
    subtasks = rp.level_subtasks(list(recipe), list(kinds))
    S = len(subtasks)
    o = types.SimpleNamespace(S=S, NA=n_agents, subtasks=subtasks, device=torch.device("cpu"),
                              names=["agent-%d" % (i + 1) for i in range(n_agents)],
                              lpairs=[], lid={}, cpairs=[], pid={})
    agsets = [(i,) for i in range(n_agents)] + ([(0, 1)] if n_agents == 2 else [])
    for s in range(S):
        for ag in agsets:
            o.lid[(s, ag)] = len(o.lpairs)
            o.lpairs.append((s, ag[0], ag[1] if len(ag) > 1 else None))
            for lvl in ((0, 1) if len(ag) == 1 and n_agents > 1 else (0,)):
                o.pid[(s, ag, lvl)] = len(o.cpairs)
                o.cpairs.append((s, ag[0], ag[1] if len(ag) > 1 else None, bool(lvl)))
    o.lpair_sub = torch.tensor([p[0] for p in o.lpairs], dtype=torch.int64)
    return o


def test_tables_cover_the_hypothesis_space():
    o = _owner()
    for model in ("bd", "up", "fb", "greedy", "dc"):
        T = ba._ObserverTables(o, 0, model)
        allocs = list(dict.fromkeys(tuple(a) for a in hypothesis_space(model, "agent-1", o.names, o.subtasks)))
        assert T.H == len(allocs) + (1 if model == "dc" else 0)
        if model in ("bd", "up", "fb"):
            assert T.H == (o.S + 1) ** 2 and int(T.static_ok.sum()) == (o.S + 1) ** 2 - 1  # only "None together" is out
        for h, alloc in enumerate(allocs):
            mine = [t for t in alloc if "agent-1" in t.subtask_agent_names]
            want = o.S if (not mine or mine[0].subtask is None) else o.subtasks.index(mine[0].subtask)
            assert int(T.sel_sub[h]) == want and bool(T.sel_joint[h]) == (bool(mine) and len(mine[0].subtask_agent_names) > 1)
            rows = [int(r) for r in T.hyp_pair[h] if int(r) != 255]
            named = [t for t in alloc if not (t.subtask is None and len(t.subtask_agent_names) > 1)
                     and (model != "greedy" or "agent-1" in t.subtask_agent_names)]
            assert len(rows) == len(named)
        assert T.P == len(T.rows) <= 128 and T.H <= 128  # what gc_bd_posterior accepts


def test_alive_mask_and_priors_match_a_python_restatement():
    o = _owner()
    T = ba._ObserverTables(o, 1, "bd")
    rng = np.random.RandomState(3)
    n, L, Pc = 64, len(o.lpairs), len(o.cpairs)
    doable = torch.from_numpy(rng.rand(n, L) < 0.6)
    inc = torch.from_numpy(rng.randint(0, 1 << o.S, size=n).astype(np.int64))
    fake = types.SimpleNamespace(lpair_sub=o.lpair_sub, cache=types.SimpleNamespace(v=None))
    alive = ba.BatchedDelegation._entries_ok(fake, T, doable, inc) & T.static_ok
    v = torch.from_numpy((1.0 + 10 * rng.rand(n, Pc)).astype(np.float32))
    v[rng.rand(n, Pc) < 0.1] = float("inf")
    fake.cache.v = v
    prior = ba.BatchedDelegation._priors(fake, T, alive, torch.arange(n))
    for e in range(n):
        weights = []
        for h, alloc in enumerate(T.allocs):
            ok = bool(T.static_ok[h])
            w = 0.0
            for t in alloc:
                if t.subtask is None:
                    continue
                s = o.subtasks.index(t.subtask)
                ag = tuple(sorted(int(nm.split("-")[1]) - 1 for nm in t.subtask_agent_names))
                ok = ok and bool(doable[e, o.lid[(s, ag)]]) and bool((int(inc[e]) >> s) & 1)
                w += 1.0 / float(v[e, o.pid[(s, ag, 0)]])
            assert ok == bool(alive[e, h]), (e, h)
            weights.append(4.0 * w if ok else 0.0)
        k = sum(1 for h in range(T.H) if bool(alive[e, h]))
        p = [wt / k for wt in weights] if k else weights
        tot = sum(p)
        want = [(1.0 / k if bool(alive[e, h]) else 0.0) if tot == 0 else p[h] / tot for h in range(T.H)]
        assert np.abs(prior[e].numpy() - np.array(want)).max() < 1e-12, e


def test_plan_cache_key_ignores_time_and_done():
    w = np.array([[5 | 9 << 6 | 17 << 24 | 1 << 31, 0x12345678, 0x1ABCDEF0, 0x00004747],
                  [5 | 9 << 6 | 99 << 24, 0x12345678, 0x1ABCDEF0, 0x00004747],
                  [6 | 9 << 6 | 17 << 24, 0x12345678, 0x1ABCDEF0, 0x00004747]], dtype=np.uint32)
    k = ba.PlanCache.key_of(torch.from_numpy(w.view(np.int32)))
    assert torch.equal(k[0], k[1]) and not torch.equal(k[0], k[2])
    back = k.contiguous().view(torch.int32).numpy().view(np.uint32)
    assert (back[:, 1:] == w[:, 1:]).all() and (back[:, 0] == (w[:, 0] & 0x00FFFFFF)).all()
