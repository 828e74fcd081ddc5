"""Differential-test harness around the UNMODIFIED reference (deletfsi/gym-cooking).

TEST INFRASTRUCTURE ONLY.  This module runs the real reference Python code from
`/root/reference/gym_cooking` (copied to a scratch dir under /tmp because the reference
uses cwd-relative paths such as `utils/levels/{}.txt`, envs/overcooked_environment.py:146,
and writes under cwd) with four tiny stub modules for the packages this image lacks
(`gym`, `termcolor`, `matplotlib.pyplot`, `pygame`; see oracle/stubs/).  It exists only in
the build container: `/root/reference` is absent on the GPU box, so nothing in the `-m gpu`
tests, `smoke()` or `bench.py` imports this file.  It is used by `oracle/gen_golden.py` to
produce the committed fixtures under `tests/golden/`, and by CPU-only tests that are
skipped when the reference is absent.

Canonical state form: agents in index order `(x, y, holding_mask)`, then
all live items sorted by `(mask, x, y, held)`, where `mask` is the 7-bit content code
    bit0 Tomato, bit1 Lettuce, bit2 Onion, bit3 Plate present;
    bit4/5/6 Tomato/Lettuce/Onion chopped
derived from `ObjectRepr.name` (= Object.full_name, utils/core.py:158-171).
"""
import argparse
import contextlib
import io
import os
import shutil
import sys
import tempfile

REF_ROOT = "/root/reference/gym_cooking"
_HERE = os.path.dirname(os.path.abspath(__file__))
_loaded = {}

KIND_BIT = {"Tomato": 0, "Lettuce": 1, "Onion": 2, "Plate": 3}


def reference_available():
    return os.path.isdir(REF_ROOT)


def load_reference():
    """Import the reference from a writable scratch copy; returns a dict of its modules."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    scratch = tempfile.mkdtemp(prefix="gc_ref_")
    dst = os.path.join(scratch, "gym_cooking")
    shutil.copytree(REF_ROOT, dst)
    sys.path.insert(0, os.path.join(_HERE, "stubs"))
    sys.path.insert(0, scratch)  # `gym_cooking.envs...` (gym_cooking/envs/__init__.py:1)
    sys.path.insert(0, dst)      # top-level `utils`, `recipe_planner`, ... as main.py sees them
    os.chdir(dst)
    with contextlib.redirect_stdout(io.StringIO()):
        import gym_cooking.envs.overcooked_environment as env_mod
        import navigation_planner.planners.e2e_brtdp as brtdp_mod
        import navigation_planner.utils as nav_utils
        import delegation_planner.bayesian_delegator as bd_mod
        import delegation_planner.utils as bd_utils
        import recipe_planner.utils as recipe_utils
        import recipe_planner.recipe as recipe_mod
        import utils.agent as agent_mod
        import utils.core as core_mod
        import utils.world as world_mod
        import utils.interact as interact_mod
    _loaded.update(dict(env=env_mod, brtdp=brtdp_mod, nav_utils=nav_utils, bd=bd_mod,
                        bd_utils=bd_utils, recipe_utils=recipe_utils, recipe=recipe_mod,
                        agent=agent_mod, core=core_mod, world=world_mod,
                        interact=interact_mod, scratch=dst))
    return _loaded


def make_arglist(level, num_agents, max_num_timesteps=100, models=None, seed=1, **kw):
    """Namespace with the flags of main.py:18-50."""
    models = list(models or [])
    models += [None] * (4 - len(models))
    ns = argparse.Namespace(
        level=level, num_agents=num_agents, max_num_timesteps=max_num_timesteps,
        max_num_subtasks=14, seed=seed, with_image_obs=False, beta=1.3, alpha=0.01, tau=2,
        cap=75, main_cap=100, play=False, record=False,
        model1=models[0], model2=models[1], model3=models[2], model4=models[3])
    for k, v in kw.items():
        setattr(ns, k, v)
    return ns


class _StubGame:
    """`step` calls self.game.get_image_obs() unconditionally (env:291) although the
    attribute only exists with --record/--with-image-obs (env:240-246)."""

    def get_image_obs(self):
        return None

    def save_image_obs(self, t):
        return None


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def make_env(level, num_agents, max_num_timesteps=100, **kw):
    ref = load_reference()
    arglist = make_arglist(level, num_agents, max_num_timesteps, **kw)
    env = ref["env"].OvercookedEnvironment(arglist)
    with quiet():
        env.reset()
    env.game = _StubGame()
    return env


def name_to_mask(full_name):
    """'Plate-ChoppedTomato' -> 7-bit content mask."""
    if full_name in (None, "None"):
        return 0
    m = 0
    for part in full_name.split("-"):
        if part == "Plate":
            m |= 1 << 3
        elif part.startswith("Fresh"):
            m |= 1 << KIND_BIT[part[5:]]
        elif part.startswith("Chopped"):
            b = KIND_BIT[part[7:]]
            m |= (1 << b) | (1 << (4 + b))
        else:
            raise ValueError(full_name)
    return m


def canonical(env):
    """(t, [(x, y, hold_mask)...], sorted [(mask, x, y, held)...]) from env.get_repr()
    (envs/overcooked_environment.py:50-62; utils/world.py:323-337)."""
    agents, items = [], []
    for entry in env.get_repr():
        fields = getattr(entry, "_fields", None)
        if fields == ("name", "location", "holding"):
            # AgentRepr (utils/agent.py:22) - or GridSquareRepr of an AgentCounter
            if entry.name.startswith("agent-"):
                agents.append((entry.location[0], entry.location[1], name_to_mask(entry.holding)))
            continue
        for o in entry:  # a name bucket: tuple of ObjectRepr / GridSquareRepr
            if getattr(o, "_fields", None) == ("name", "location", "is_held"):
                items.append((name_to_mask(o.name), o.location[0], o.location[1], int(bool(o.is_held))))
    items.sort()
    return env.t, agents, items


ACTIONS = [(0, 1), (0, -1), (-1, 0), (1, 0), (0, 0)]  # world.py:16 order + stay


def run_trace(level, num_agents, actions, max_num_timesteps=100):
    """Replay `actions[T][n]` (indices into ACTIONS) through the reference env.step.

    Returns a list of per-step dicts: canonical state after the step, reward, done,
    number of CollisionRepr appended this step, and the post-collision executed actions.
    Stops at `done` (the reference main loop stops there, main.py:97) or when the reference
    crashes on the co-located-holders assertion (SURVEY.md §7 "Hard parts")."""
    env = make_env(level, num_agents, max_num_timesteps)
    out = [dict(state=canonical(env), reward=0, done=False, ncoll=0, crashed=False)]
    names = env.get_agent_names()
    for step_actions in actions:
        ad = {names[i]: ACTIONS[a] for i, a in enumerate(step_actions)}
        ncoll0 = len(env.collisions)
        try:
            with quiet():
                _, reward, done, info = env.step(ad)
        except (AssertionError, AttributeError):  # world.py:417 (its message itself raises)
            out.append(dict(crashed=True))
            break
        executed = [ACTIONS.index(tuple(env.agent_actions[nm])) for nm in names]
        out.append(dict(state=canonical(env), reward=reward, done=bool(done),
                        ncoll=len(env.collisions) - ncoll0, executed=executed,
                        termination_info=info["termination_info"], crashed=False))
        if done:
            break
    return out
