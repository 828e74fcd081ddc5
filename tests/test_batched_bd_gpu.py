"""The device-resident Bayesian-Delegation loop (batched_agents.BatchedDelegation) against the host
facade loop (main.main_loop over RealAgent / BayesianDelegator / E2E_BRTDP): with every random
tie-break made deterministic in BOTH, each env of a batch must pick the facade's actions step by
step; with random tie-breaks the batch must finish like the reference's episodes do."""
import argparse

import numpy as np
import pytest
import torch

import gym_cooking_b200 as gcb
from gym_cooking_b200 import batched_agents, delegation_planner, main as gmain, navigation_planner

pytestmark = pytest.mark.gpu


def _arglist(level, models, max_t=100):
    models = list(models) + [None] * (4 - len(models))
    return argparse.Namespace(level=level, num_agents=sum(m is not None for m in models), max_num_timesteps=max_t,
                              max_num_subtasks=14, seed=1, beta=1.3, alpha=0.01, tau=2, cap=75, main_cap=100,
                              play=False, record=False, with_image_obs=False, model1=models[0], model2=models[1],
                              model3=models[2], model4=models[3])


@pytest.fixture
def deterministic_facade(monkeypatch):
    """first-minimum argmin, canonical-key arg-max, 'do nothing' = stay"""
    subtasks = {}

    def get_max(self):
        if not self.probs:
            return None
        best = max(self.probs.values())
        cands = [a for a, p in self.probs.items() if p >= best - 1e-12]
        return min(cands, key=lambda a: batched_agents.alloc_key(a, subtasks["list"]))

    monkeypatch.setattr(delegation_planner.SubtaskAllocDistribution, "get_max", get_max)
    monkeypatch.setattr(navigation_planner, "argmin", lambda v: int(np.argmin(np.asarray(v, dtype=np.float64))))
    monkeypatch.setattr(np.random, "choice", lambda n, p=None: n - 1)
    return subtasks


@pytest.mark.parametrize("level,models", [("open-divider_tomato", ("bd", "bd")),
                                          ("partial-divider_tomato", ("bd", "up")),
                                          ("open-divider_tl", ("bd", "bd")),
                                          ("partial-divider_tl", ("greedy", "bd")),
                                          ("open-divider_tomato", ("dc", "fb")),
                                          # three and four agents: joint rows the observer is not part of (25-entry
                                          # likelihood rows), level-1 joint planning, every pair of agents
                                          ("open-divider_tomato", ("bd", "bd", "bd")),
                                          ("partial-divider_tomato", ("bd", "up", "dc")),
                                          ("open-divider_tl", ("greedy", "bd", "fb")),
                                          ("open-divider_tomato", ("bd", "up", "dc", "greedy")),
                                          ("full-divider_salad", ("bd", "dc", "bd")),  # 2 457-row table: lists
                                          ("partial-divider_tl", ("up", "bd", "greedy", "dc"))])  # 8 028 rows
def test_batched_loop_equals_facade_loop(level, models, deterministic_facade):
    _compare_with_facade(level, models, deterministic_facade)


@pytest.mark.parametrize("level,models", [("open-divider_tl", ("bd", "bd")), ("partial-divider_tl", ("greedy", "bd")),
                                          ("open-divider_tomato", ("dc", "fb")), ("partial-divider_tomato", ("bd", "up", "dc"))])
def test_list_form_equals_facade_loop(level, models, deterministic_facade, monkeypatch):
    """the same comparison with the small hypothesis tables forced through the per-env LIST form that large tables
    (three agents on tl / salad levels, four agents) always take"""
    monkeypatch.setattr(batched_agents, "_FORCE_LISTS", True)
    _compare_with_facade(level, models, deterministic_facade)


def _compare_with_facade(level, models, deterministic_facade):
    loop = batched_agents.BatchedDelegation(level, 8, models, deterministic=True)
    assert all(T.lists == (T.H > batched_agents.DENSE_MAX_H or batched_agents._FORCE_LISTS) for T in loop.tables)
    deterministic_facade["list"] = loop.subtasks
    env, agents, history = gmain.main_loop(_arglist(level, models), max_steps=40)
    for step, action_dict in enumerate(history):
        loop.step()
        want = [gcb.ACTION_INDEX[tuple(action_dict["agent-%d" % (i + 1)])] for i in range(len(models))]
        got = loop.last_actions.cpu().numpy()
        assert (got == np.array(want, dtype=np.uint8)[None, :]).all(), "step %d: facade %s, batched %s" % (
            step, want, got[0].tolist())
    words = np.array([int(w) & 0xFFFFFFFF for w in env.state[0].tolist()], dtype=np.uint32)
    assert (loop.kb.state.cpu().numpy().view(np.uint32) == words[None, :]).all()
    assert bool(loop.kb.done.all()) == bool(env.done())


@pytest.mark.parametrize("level,models,limit,n", [("open-divider_tomato", ("bd", "bd"), 40, 512),
                                                  ("partial-divider_tl", ("bd", "bd"), 100, 512),
                                                  ("open-divider_tomato", ("bd", "bd"), 40, 8192)])  # compacts on the way
def test_random_tie_breaks_finish(level, models, limit, n):
    loop = batched_agents.BatchedDelegation(level, n, models, seed=3)
    steps = loop.run()
    stats = loop.kb.stats().cpu().numpy()
    done = loop.kb.done
    assert bool(done.all())
    success = int(stats[1])
    t = ((loop.kb.state[:, 0].to(torch.int64) >> 24) & 127).float()
    print("%s %s: %d/%d delivered, mean %.1f steps (%d loop steps), %d planning states solved for %d lookups" % (
        level, models, success, n, float(t.mean()), steps, loop.cache.solved_states, loop.cache.lookups))
    assert success >= 0.9 * n
    assert float(t[loop.kb.reward.bool()].mean()) <= limit
    assert loop.cache.solved_states < loop.cache.lookups / 4  # the memo is doing its job
    assert int(stats[0]) == n and loop.kb.num_envs == n and (n < 4096 or loop.wkb.num_envs < n)


@pytest.mark.parametrize("level,models", [("open-divider_salad", ("bd", "bd")), ("partial-divider_tl", ("greedy", "dc"))])
def test_likelihood_row_kernel_equals_the_torch_restatement(level, models):
    """gc_bd_likelihood_rows (prob_nav_actions' softmax inputs, bd:461-689) against the tensor-op
    restatement, on the states / executed actions of a running delegation loop"""
    loop = batched_agents.BatchedDelegation(level, 2048, models, seed=5)
    compared = 0
    for step in range(14):
        loop.step()
        if step % 3 != 1:
            continue
        for T in loop.tables:
            qd, nv, ai = loop._likelihood_rows(T)
            qd_t, nv_t, ai_t = loop._likelihood_rows_torch(T)
            assert torch.equal(nv, nv_t) and torch.equal(ai, ai_t)
            live = torch.arange(5, device=qd.device)[None, None, :] < nv[:, :, None]
            assert torch.equal(torch.where(live, qd, 0.0), torch.where(live, qd_t, 0.0))
            assert bool((nv >= 1).all()) and bool((ai < nv).all())
            compared += int(nv.numel())
    assert compared > 100000


@pytest.mark.parametrize("level,optimum,min_success", [("open-divider_tomato", 15, 0.99), ("partial-divider_tomato", 17, 0.9),
                                                       ("open-divider_salad", 24, 0.5)])
def test_episode_lengths_against_the_reference_yardsticks(level, optimum, min_success):
    """Distributional sanity of whole bd/bd episodes: no delivered episode is shorter than the
    optimal length the reference's plots use as yardstick (misc/metrics/make_graphs.py:48-52: open tomato 15,
    partial tomato 17, open salad 24 steps for two agents), the mean stays within 2x of it, and the reference's
    own open-divider_tomato bd/bd run (seed 1: delivered at t = 23, profiles/r02_python_reference_cpu.json) lies
    inside the batch's distribution."""
    n = 4096
    loop = batched_agents.BatchedDelegation(level, n, ("bd", "bd"), seed=11)
    loop.run()
    t = ((loop.kb.state[:, 0].to(torch.int64) >> 24) & 127)
    ok = loop.kb.reward.bool()
    assert float(ok.float().mean()) >= min_success
    lengths = t[ok]
    assert int(lengths.min()) >= optimum, (int(lengths.min()), optimum)
    assert float(lengths.float().mean()) <= 2.0 * optimum
    if level == "open-divider_tomato":
        assert int(lengths.min()) <= 23 <= int(lengths.max())
    print("%s: %.1f %% delivered, steps min %d mean %.1f max %d (reference yardstick %d)" % (
        level, 100 * float(ok.float().mean()), int(lengths.min()), float(lengths.float().mean()), int(lengths.max()), optimum))


def test_mixed_four_agent_grid_runs():
    """cfg-5 in miniature: four agents, rotated model types, three levels - every episode ends (delivery or
    horizon), most deliver, and the totals add up"""
    levels = ("open-divider_tomato", "partial-divider_tl", "full-divider_salad")
    r = batched_agents.run_mixed(24, 4, levels=levels, horizon=60, seed=3)
    assert r["envs"] == 72 and len(r["per_level"]) == 3
    assert [rec["models"] for rec in r["per_level"]] == ["bd/up/dc/fb", "up/dc/fb/greedy", "dc/fb/greedy/bd"]
    assert r["agent_steps"] == sum(rec["agent_steps"] for rec in r["per_level"]) and r["agent_steps"] <= 72 * 4 * 61
    assert r["delivered"] >= 36 and r["posterior_updates"] > 0
    assert r["per_level"][2]["hypotheses"][2] == 10 and max(r["per_level"][2]["hypotheses"]) == 39906  # greedy / bd tables
    assert all(rec["loop_steps"] <= 61 for rec in r["per_level"])


@pytest.mark.parametrize("level,models", [("open-divider_salad", ("bd", "bd")), ("partial-divider_tl", ("bd", "up", "dc", "greedy")),
                                          ("full-divider_tomato", ("greedy", "bd", "fb"))])
def test_agent_view_kernels_equal_their_torch_twins(level, models):
    """gc_offered_actions / gc_subtasks_completed (what a RealAgent reads off the env each step) against the tensor-op
    restatements `single_actions` / `goal_count`, on the states of a running loop"""
    from gym_cooking_b200 import planning
    loop = batched_agents.BatchedDelegation(level, 1024, models, seed=9)
    offered_rows = completed = 0
    for step in range(30):
        before = loop.kb.state.clone()
        loop.step()
        kb = loop.kb
        bits = planning.offered_actions(kb)
        want = loop.single_actions(kb.state)  # bool[N][NA][4]
        got = ((bits[:, :, None] >> torch.arange(4, dtype=torch.uint8, device=bits.device)[None, None, :]) & 1).bool()
        assert torch.equal(got, want), step
        offered_rows += int(want.numel())
        # every subtask index for every agent, not only the ones the agents happen to follow
        for s in range(loop.S + 1):
            sub = torch.full((kb.num_envs, loop.NA), s, dtype=torch.uint8, device=kb.device)
            done = planning.subtasks_completed(kb, before, sub)
            if s < loop.S:
                idx = torch.full((kb.num_envs,), s, dtype=torch.int64, device=kb.device)
                ref = loop.goal_count(kb.state, idx) > loop.goal_count(before, idx)
            else:
                ref = torch.zeros(kb.num_envs, dtype=torch.bool, device=kb.device)
            assert torch.equal(done.bool(), ref[:, None].expand(-1, loop.NA)), (step, s)
            completed += int(ref.sum())
    assert offered_rows > 0 and completed > 100
