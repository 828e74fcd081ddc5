"""Hypothesis space of the BayesianDelegator facade against the reference's own
(add_subtasks / add_dc_subtasks / add_greedy_subtasks, bayesian_delegator.py:792-1000)."""
import pytest

from gym_cooking_b200 import delegation_planner as dp, recipe_planner as rp

OBJECTS = ["Tomato", "Lettuce", "Plate", "Plate"]


def _canon(allocs):
    return sorted(set("|".join(sorted("%s:%s" % (t.subtask, ",".join(t.subtask_agent_names)) for t in a)) for a in allocs))


def test_sizes_match_survey():
    """raw allocations 16 (2 agents, tomato... salad) -> 39 906 (4 agents, salad)."""
    salad = rp.level_subtasks(["Salad"], OBJECTS)
    two = dp.hypothesis_space("bd", "agent-1", ["agent-1", "agent-2"], salad[:3])
    assert len(set(map(tuple, two))) == 16
    four = dp.hypothesis_space("bd", "agent-1", ["agent-%d" % i for i in range(1, 5)], salad)
    assert len(set(map(tuple, four))) == 39906
    assert len(dp.hypothesis_space("greedy", "agent-2", ["agent-1", "agent-2"], salad)) == len(salad) + 1


def test_distribution_helpers():
    d = dp.SubtaskAllocDistribution([[dp.SubtaskAllocation(None, ("agent-1",))], [dp.SubtaskAllocation(None, ("agent-2",))]])
    assert list(d.probs.values()) == [0.5, 0.5]
    a, b = d.enumerate_subtask_allocs()
    d.update(a, 3.0)
    d.normalize()
    assert abs(d.get(a) - 0.75) < 1e-12 and d.get_max() == a
    d.set(a, 0.0), d.set(b, 0.0)
    d.normalize()
    assert d.get(a) == 0.5
    d.delete(a)
    assert d.enumerate_subtask_allocs() == [b]


@pytest.mark.parametrize("model", ["bd", "dc", "greedy"])
@pytest.mark.parametrize("n_agents", [1, 2, 3])
def test_matches_reference_when_present(model, n_agents):
    import ref_harness as H
    if not H.reference_available():
        pytest.skip("reference not mounted")
    ref = H.load_reference()
    env = H.make_env("open-divider_salad", n_agents)
    names = env.get_agent_names()
    for n_sub in (2, 4):
        subs = list(env.all_subtasks)[:n_sub]
        with H.quiet():
            d = ref["bd"].BayesianDelegator(agent_name=names[0], all_agent_names=names, model_type=model,
                                            planner=None, none_action_prob=0.5)
            d.incomplete_subtasks = list(subs)
            expect = d.get_subtask_alloc_probs().enumerate_subtask_allocs()
        mine_subs = [getattr(rp, type(s).__name__)(*s.args) for s in subs]
        got = dp.hypothesis_space(model, names[0], names, mine_subs)
        assert _canon(got) == _canon(expect), (model, n_agents, n_sub)
