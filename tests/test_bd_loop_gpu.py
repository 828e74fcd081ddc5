"""The reference's main loop (main.py:85-117) on the facades: RealAgent -> BayesianDelegator ->
E2E_BRTDP -> OvercookedEnvironment, every number from the kernels.  The reference needs 23 steps /
362 s for this episode with bd/bd; the optimum is 15 (make_graphs.py:49)."""
import argparse
import time

import pytest

import gym_cooking_b200 as gcb
from gym_cooking_b200 import main as gmain

pytestmark = pytest.mark.gpu


def _arglist(level, n_agents, models, max_t=100):
    models = list(models) + [None] * (4 - len(models))
    return argparse.Namespace(level=level, num_agents=n_agents, max_num_timesteps=max_t, max_num_subtasks=14, seed=1,
                              beta=1.3, alpha=0.01, tau=2, cap=75, main_cap=100, play=False, record=False,
                              with_image_obs=False, model1=models[0], model2=models[1], model3=models[2], model4=models[3])


@pytest.mark.parametrize("level,models,limit", [("open-divider_tomato", ("bd", "bd"), 30),
                                                ("open-divider_tomato", ("greedy", "greedy"), 40),
                                                ("partial-divider_tomato", ("dc", "dc"), 45),
                                                ("open-divider_tl", ("up", "bd"), 60)])
def test_episode_completes(level, models, limit):
    gmain.fix_seed(1)
    t0 = time.time()
    env, agents, history = gmain.main_loop(_arglist(level, len(models), models))
    assert env.successful, (env.termination_info, env.t)
    assert env.t <= limit, env.t
    assert env.termination_info == "Terminating because all deliveries were completed"
    assert all(a.all_done() for a in agents) or env.successful
    print("%s %s: success in %d steps, %.1f s" % (level, models, env.t, time.time() - t0))


def test_delegator_and_planner_surface():
    env = gcb.make(arglist=_arglist("open-divider_tomato", 2, ("bd", "bd")))
    obs = env.reset()
    from gym_cooking_b200.navigation_planner import E2E_BRTDP, PlannerLevel
    from gym_cooking_b200.delegation_planner import BayesianDelegator
    planner = E2E_BRTDP(alpha=0.01, tau=2, cap=75, main_cap=100)
    chop = obs.all_subtasks[0]
    a = planner.get_next_action(env=obs, subtask=chop, subtask_agent_names=("agent-2",), other_agent_planners={})
    key = (planner.cur_state.get_repr(), chop)
    assert abs(planner.v_l[key] - 8.8) < 1e-5 and planner.v_l[key] == planner.v_u[key]  # SURVEY section 8a KAT
    assert a in planner.get_actions() and abs(planner.Q(obs, a) - 8.8) < 1e-5
    assert planner.planner_level == PlannerLevel.LEVEL0 and not planner.is_joint
    joint = planner.get_next_action(env=obs, subtask=chop, subtask_agent_names=("agent-1", "agent-2"), other_agent_planners={})
    assert planner.is_joint and len(joint) == 2 and abs(planner.v_l[key] - 7.8) < 1e-5
    d = BayesianDelegator("agent-1", obs.get_agent_names(), "bd", planner, 0.5)
    assert d.should_reset_priors(obs, list(obs.all_subtasks))
    d.set_priors(obs, list(obs.all_subtasks), "spatial")
    allocs = d.probs.enumerate_subtask_allocs()
    # at reset only Chop(Tomato) is doable: [Chop(1,2)], [Chop(1);None(2)], [None(1);Chop(2)]
    assert len(allocs) == 3 and abs(sum(d.probs.probs.values()) - 1) < 1e-12
    pri = {tuple((str(t.subtask), t.subtask_agent_names) for t in a): p for a, p in d.probs.get_list()}
    w = {"joint": 1 / 7.8, "a1": 1 / 13.2, "a2": 1 / 8.8}
    z = sum(w.values())
    assert abs(pri[(("Chop(Tomato)", ("agent-1", "agent-2")),)] - w["joint"] / z) < 1e-6
    obs2, _, _, _ = env.step({"agent-1": (0, 1), "agent-2": (1, 0)})
    before = dict(d.probs.probs)
    d.bayes_update(env.obs_tm1, env.agent_actions, 1.3)
    assert abs(sum(d.probs.probs.values()) - 1) < 1e-12 and d.probs.probs != before
    st, names = d.select_subtask("agent-1")
    assert "agent-1" in names


def test_bag_and_record(tmp_path, monkeypatch):
    """main_loop writes the reference's Bag pickle and, with --record, one PNG per step (+ t=000)"""
    import pickle
    from gym_cooking_b200.misc.game.gameimage import decode_png
    monkeypatch.chdir(tmp_path)
    gmain.fix_seed(1)
    args = _arglist("open-divider_tomato", 2, ("bd", "bd"))
    args.record = True
    env, agents, history = gmain.main_loop(args, bag_directory=str(tmp_path / "pickles"))
    data = pickle.load(open(env.bag_path, "rb"))
    assert data["was_successful"] and data["num_completed_subtasks_end"] == 3 and data["num_total_subtasks"] == 3
    assert len(data["actions"]["agent-1"]) == len(history) == env.t
    assert data["actions"]["agent-2"][0] == history[0]["agent-2"]
    frames = sorted((tmp_path / "misc" / "game" / "record" / env.filename).iterdir())
    assert [f.name for f in frames] == ["t=%03d.png" % t for t in range(env.t + 1)]
    first, last = decode_png(frames[0].read_bytes()), decode_png(frames[-1].read_bytes())
    assert first.shape == (560, 560, 3) and (first != last).any()
