"""The reference-shaped facade (OvercookedEnvironment / obs / info) on the GPU path."""
import argparse
import os

import numpy as np
import pytest
import torch

import gym_cooking_b200 as gcb
from gym_cooking_b200.utils.core import name_to_mask

pytestmark = pytest.mark.gpu


def _arglist(level, n_agents, max_t=100):
    return argparse.Namespace(level=level, num_agents=n_agents, max_num_timesteps=max_t, max_num_subtasks=14,
                              seed=1, model1="bd", model2="bd", model3=None, model4=None, play=False, record=False,
                              with_image_obs=False)


def _canonical(obs):
    """same reduction the golden generator applies to the reference's env.get_repr()"""
    agents, items = [], []
    for entry in obs.get_repr():
        if getattr(entry, "_fields", None) == ("name", "location", "holding"):
            agents.append((entry.location[1] * 8 + entry.location[0], name_to_mask(entry.holding)))
            continue
        for o in entry:
            if getattr(o, "_fields", None) == ("name", "location", "is_held"):
                items.append((name_to_mask(o.name) << 7) | ((o.location[1] * 8 + o.location[0]) << 1) | int(o.is_held))
    return agents, sorted(items)


def test_single_env_drop_in_follows_golden_trace(golden_dir):
    tr = np.load(os.path.join(golden_dir, "env_traces.npz"))
    rows = [r for r in range(tr["meta"].shape[0]) if tr["meta"][r, 1] in (2, 3) and tr["meta"][r, 3] == 0][:6]
    rows += [next(r for r in range(tr["meta"].shape[0]) if tr["reward"][r, tr["length"][r]] == 1)]
    for r in rows:
        lvl, n_agents, max_t, _ = (int(v) for v in tr["meta"][r])
        env = gcb.make("gym_cooking:overcookedEnv-v0", arglist=_arglist(str(tr["levels"][lvl]), n_agents, max_t))
        obs = env.reset()
        assert obs.t == 0 and len(obs.sim_agents) == n_agents
        names = env.get_agent_names()
        ncoll = 0
        for s in range(int(tr["length"][r])):
            ad = {names[i]: gcb.ACTIONS[int(tr["actions"][r, s, i])] for i in range(n_agents)}
            state_before = _canonical(env)
            obs, reward, done, info = env.step(ad)
            assert set(info) == {"t", "obs", "image_obs", "done", "termination_info"}
            agents, items = _canonical(obs)
            assert agents == [tuple(int(v) for v in a) for a in tr["agents"][r, s + 1, :n_agents]]
            assert items == [int(k) for k in tr["keys"][r, s + 1] if k != 0x3FFF]
            assert info["t"] == env.t == s + 1 == obs.t
            assert done == bool(tr["done"][r, s + 1]) == env.done()
            assert reward == int(tr["reward"][r, s + 1]) == env.reward()
            assert [gcb.ACTION_INDEX[env.agent_actions[nm]] for nm in names] == \
                [int(v) for v in tr["executed"][r, s + 1, :n_agents]]
            assert _canonical(env.obs_tm1) == state_before  # obs_tm1 = state before, executed actions (env:273)
            assert [gcb.ACTION_INDEX[a.action] for a in env.obs_tm1.sim_agents] == \
                [int(v) for v in tr["executed"][r, s + 1, :n_agents]]
            ncoll += int(tr["ncoll"][r, s + 1])
            assert len(env.collisions) == ncoll
        if tr["done"][r, tr["length"][r]]:
            expect = ("Terminating because all deliveries were completed" if env.successful
                      else "Terminating because passed %d timesteps" % max_t)
            assert env.termination_info == expect


def test_batched_facade_host_actions():
    n = 4096
    env = gcb.OvercookedEnvironment(_arglist("partial-divider_tl", 2), num_envs=n)
    obs = env.reset()
    ref = gcb.KitchenBatch("partial-divider_tl", 2, n, 100)
    acts = ref.random_actions(30, seed=9)
    for s in range(30):
        host = acts[s].cpu().pin_memory()
        obs, reward, done, info = env.step(host)
        ref.step(acts[s])
        assert reward.shape == (n,) and done.tensor().dtype == torch.bool and not done.tensor().is_cuda
    assert torch.equal(env.state, ref.state)
    assert torch.equal(done.tensor(), ref.done.cpu()) and reward.sum() == int(ref.reward.sum())
    # large batches take the chunk-pipelined path (H2D / kernel / D2H of different chunks overlap)
    big, bref = gcb.OvercookedEnvironment(_arglist("partial-divider_tl", 2, 25), num_envs=1 << 17), \
        gcb.KitchenBatch("partial-divider_tl", 2, 1 << 17, 25)
    big.reset()
    bacts = bref.random_actions(30, seed=3)
    for s in range(30):
        big.PIPELINE_CHUNKS = 4 if s % 2 else 1  # alternate the opt-in multi-stream path with the default
        _, breward, bdone, _ = big.step(bacts[s].cpu().pin_memory())
        bref.step(bacts[s])
    assert torch.equal(big.state, bref.state) and bdone.all() and breward.sum() == 0
    assert torch.equal(bdone.tensor(), bref.done.cpu())
    view = obs[17]
    w = ref.state[17].tolist()
    assert view.t == 30 and [a.location for a in view.sim_agents] == \
        [(x, y) for (x, y, _) in gcb.decode_state(w, 2)["agents"]]
    assert len(env.all_subtasks) == 6 and str(env.all_subtasks[0]) == "Chop(Tomato)"


def test_reference_named_helpers():
    """display / lower-bound / recipe helpers the reference's callers use on env and obs (env:66-98, 378-473, 594-664)"""
    import argparse
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=100, max_num_subtasks=14, seed=1,
                            model1="bd", model2="bd", model3=None, model4=None, record=False, with_image_obs=False)
    env = gcb.make(arglist=ns)
    obs = env.reset()
    text = str(obs)
    rows = text.split("\n")
    assert len(rows) == 7 and all(len(r) >= 14 for r in rows)
    assert "1" in text and "2" in text and "t" in text and "p" in text and "/" in text and "*" in text
    assert str(env) == text
    assert [str(s) for s in obs.run_recipes()] == [str(s) for s in env.all_subtasks]
    chop = env.all_subtasks[0]
    lb1 = obs.get_lower_bound_for_subtask_given_objs(chop, ("agent-1",), None, None, None)
    lb2 = env.get_lower_bound_for_subtask_given_objs(chop, ("agent-1", "agent-2"), None, None, None)
    assert 1.0 <= lb2 <= lb1 < obs.world.perimeter
    from gym_cooking_b200.utils.agent import RealAgent
    agent = RealAgent(ns, "agent-1", "blue", env.recipes)
    assert [str(s) for s in agent.get_subtasks(obs.world)] == [str(s) for s in env.all_subtasks]
    agent.setup_subtasks(env=obs)
    assert agent.delegator.add_subtasks() and str(agent) == "1"


def test_step_async_wait_equals_step():
    """Vector-env style step_async / step_wait on two alternating batches (own streams) returns what the
    synchronous step returns, step by step."""
    import argparse
    n = 40000
    ns = argparse.Namespace(level="partial-divider_tl", num_agents=2, max_num_timesteps=12, max_num_subtasks=14, seed=1,
                            model1=None, model2=None, model3=None, model4=None)
    envs = [gcb.OvercookedEnvironment(ns, num_envs=n, track_collisions=False) for _ in range(3)]
    for e in envs:
        e.reset()
    a, b, ref = envs
    acts = ref._kb.random_actions(14, seed=4)
    joint = [(acts[s][:, 0] * 5 + acts[s][:, 1]).to(torch.uint8).cpu().pin_memory() for s in range(14)]
    for s in range(14):
        a.step_async(joint[s])
        b.step_async(joint[s])
        _, r_ref, d_ref, _ = ref.step(joint[s])
        _, r_a, d_a, info = a.step_wait()
        _, r_b, d_b, _ = b.step_wait()
        assert torch.equal(a.state, ref.state) and torch.equal(b.state, ref.state), s
        for got in ((r_a, d_a), (r_b, d_b)):
            assert torch.equal(got[0].tensor().bool(), r_ref.tensor().bool()), s
            assert torch.equal(got[1].tensor().bool(), d_ref.tensor().bool()), s
        assert info["t"] == s + 1
    assert bool(d_ref.tensor().all())
