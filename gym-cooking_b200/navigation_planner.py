"""`E2E_BRTDP` with the reference's call signatures, answered by the planner kernels.

Reference: navigation_planner/planners/e2e_brtdp.py - ctor(alpha, tau, cap, main_cap) :45,
set_settings(env, subtask, subtask_agent_names, other_agent_planners={}) :582,
get_next_action(env, subtask, subtask_agent_names, other_agent_planners) :987, Q(state, action,
value_f) :740, get_actions(state_repr) :151, attributes v_l / v_u (dicts keyed (repr, subtask)),
cur_state, start, is_joint, subtask, subtask_agent_names, goal_obj, is_subtask_complete,
planner_level.

The reference brackets the optimal value of the (subtask, agent set) MDP with BRTDP and its answer
depends on exploration order, RNG tie-breaks and caps.  Here `set_settings` asks gc_subtask_q /
gc_joint_q for the fixed point itself, so v_l == v_u == V* and Q(s, a) = cost + V*(T(s, a)) for the
planning start state; alpha / tau / cap / main_cap are accepted and unused.  Planning level: an
empty `other_agent_planners` is level 0 (others frozen into counters), a non-empty one level 1
(others are plain obstacles); as in the reference, the other agents' predicted actions are never
applied to the transition model (e2e:974-977 vs :124-140).
"""
from enum import Enum

import numpy as np
import torch

from . import engine, planning, recipe_planner
from .utils.core import Object

_INF = float("inf")


class PlannerLevel(Enum):
    LEVEL1 = 1
    LEVEL0 = 0


def argmin(vector):
    """Uniform random choice among the minima (e2e_brtdp.py:27-30 always draws from np.random)."""
    v = np.asarray(vector, dtype=np.float64)
    e_x = v == v.min()
    return int(np.where(np.random.multinomial(1, e_x / e_x.sum()))[0][0])


_batches = {}


def _plan_batch(env):
    """A one-env KitchenBatch per (level, n_agents, device), reused as the kernels' state holder."""
    owner = getattr(env, "_env", env)
    kb = owner.batch
    key = (tuple(kb.level_names), kb.num_agents, str(kb.device))
    if key not in _batches:
        _batches[key] = engine.KitchenBatch(kb.level_names[0], kb.num_agents, 1, kb.max_num_timesteps, device=kb.device)
    return _batches[key]


def packed_words(env):
    """packed state of an OvercookedEnvironment facade or one of its obs views"""
    if hasattr(env, "_words"):
        return list(env._words)
    return [int(w) & 0xFFFFFFFF for w in env.state[0].tolist()]


def solve_pair(env, subtask, subtask_agent_names, level1=False):
    """(V*, {action: Q}, status) of one (subtask, agent set) at env's state.  Actions are (dx, dy)
    tuples, or pairs of them for two agents (in agent-name order, e2e:661)."""
    pb = _plan_batch(env)
    words = np.array(packed_words(env), dtype=np.uint32).view(np.int32)
    pb.state.copy_(torch.from_numpy(words).to(pb.device).view(1, 4))
    pb.set_subtask_masks([recipe_planner.subtask_masks(subtask)])
    idx = sorted(int(nm.split("-")[1]) - 1 for nm in subtask_agent_names)
    pair = (0, idx[0], idx[1] if len(idx) > 1 else None, bool(level1))
    v, q, status = planning.subtask_q(pb, [pair])
    v, q, status = float(v[0, 0]), q[0, 0].tolist(), int(status[0, 0])
    table = {}  # NaN = not offered by get_single_actions / is_collision; +inf = offered, goal out of reach
    if len(idx) == 1:
        for a in range(5):
            if q[a] == q[a]:
                table[engine.ACTIONS[a]] = q[a]
    else:
        for a in range(25):
            if q[a] == q[a]:
                table[(engine.ACTIONS[a // 5], engine.ACTIONS[a % 5])] = q[a]
    return v, table, status


def lower_bound_pair(env, subtask, subtask_agent_names):
    """env.get_lower_bound_for_subtask_given_objs (env:594-664) of one (subtask, agent set) at env's state"""
    pb = _plan_batch(env)
    words = np.array(packed_words(env), dtype=np.uint32).view(np.int32)
    pb.state.copy_(torch.from_numpy(words).to(pb.device).view(1, 4))
    pb.set_subtask_masks([recipe_planner.subtask_masks(subtask)])
    idx = sorted(int(nm.split("-")[1]) - 1 for nm in subtask_agent_names)
    return float(planning.lower_bound(pb, [(0, idx[0], idx[1] if len(idx) > 1 else None)])[0, 0])


def goal_count(world, subtask, goal_obj, delivery_locs):
    """e2e_brtdp._define_goal_state :435-566: how many goal objects the world holds"""
    if isinstance(subtask, recipe_planner.Deliver):
        return len([o for o in world.get_object_locs(goal_obj, is_held=False) if o in delivery_locs])
    return len(world.get_all_object_locs(goal_obj))


class E2E_BRTDP:
    """Navigation planner facade (exact values instead of BRTDP bounds)."""

    def __init__(self, alpha, tau, cap, main_cap):
        self.alpha, self.tau, self.cap, self.main_cap = alpha, tau, cap, main_cap
        self.v_l, self.v_u = {}, {}
        self.start = self.cur_state = None
        self.is_joint = False
        self.planner_level = PlannerLevel.LEVEL0
        self.subtask = None
        self.subtask_agent_names = ()
        self.goal_obj = self.start_obj = None
        self.removed_object = None
        self.other_agent_planners = {}
        self.is_subtask_complete = lambda w: False
        self.time_cost, self.action_cost = 1.0, 0.1
        self._q, self._status = {}, 0

    def __copy__(self):
        c = E2E_BRTDP(self.alpha, self.tau, self.cap, self.main_cap)
        c.__dict__ = self.__dict__.copy()  # value dicts are shared with the copy, as in the reference (:96-101)
        return c

    # -- configuration ---------------------------------------------------------------------
    def set_settings(self, env, subtask, subtask_agent_names, other_agent_planners={}):
        assert len(subtask_agent_names) <= 2, "Cannot have more than 2 agents!"
        self.planner_level = PlannerLevel.LEVEL1 if other_agent_planners else PlannerLevel.LEVEL0
        self.other_agent_planners = other_agent_planners or {}
        self.subtask = subtask
        self.subtask_agent_names = tuple(subtask_agent_names)
        self.is_joint = len(subtask_agent_names) == 2
        self.start = self.cur_state = env
        self.removed_object = None
        if subtask is None:
            self.start_obj = self.goal_obj = None
            self.is_goal_state = lambda h: True
            self._q, self._status = {((0, 0), (0, 0)) if self.is_joint else (0, 0): 0.0}, 1
            value = 0.0
        else:
            kind, a, b, goal = recipe_planner.subtask_masks(subtask)
            self.goal_obj = Object((None, None), goal)
            self.start_obj = ([Object((None, None), a), Object((None, None), b)] if kind == recipe_planner.ST_MERGE
                              else Object((None, None), a))
            delivery = [gs.location for gs in env.world.objects.get("Delivery", [])]
            base = goal_count(env.world, subtask, self.goal_obj, delivery)
            if self.planner_level == PlannerLevel.LEVEL0:  # an object held by a frozen agent is deleted (:397-399)
                for ag in env.sim_agents:
                    if ag.name not in subtask_agent_names and ag.holding is not None:
                        self.removed_object = ag.holding
            extra = 1 if (self.removed_object is not None and self.removed_object == self.goal_obj) else 0
            self.cur_obj_count = base
            self.is_subtask_complete = lambda w: goal_count(w, subtask, self.goal_obj, delivery) + extra > base
            value, self._q, self._status = solve_pair(env, subtask, subtask_agent_names,
                                                      self.planner_level == PlannerLevel.LEVEL1)
        key = (env.get_repr(), subtask)
        self.v_l[key] = self.v_u[key] = value

    # -- queries ---------------------------------------------------------------------------
    def get_actions(self, state_repr=None):
        """valid (joint) actions at the planning start state (:151-206)"""
        if self.subtask is None:
            return [(0, 0)]
        return list(self._q.keys())

    def Q(self, state, action, value_f=None):
        """cost(s, a) + V*(T(s, a)) at the planning start state (:740-779); +inf for an invalid action"""
        return self._q.get(tuple(action) if not self.is_joint else (tuple(action[0]), tuple(action[1])), _INF)

    def V(self, state, _type):
        """V(s) = min_a Q(s, a) at the planning start state (:784-809); lower == upper here"""
        return min(self._q.values()) if self._q else 0.0

    def value_init(self, env_state):  # :701-737 - the bounds are exact from the start
        key = (env_state.get_repr(), self.subtask)
        self.v_l.setdefault(key, self.v_l.get((self.start.get_repr(), self.subtask), 0.0))
        self.v_u.setdefault(key, self.v_l[key])

    def get_subtask_agents(self, env_state):  # :645-659
        return [a for a in env_state.sim_agents if a.name in self.subtask_agent_names]

    def cost(self, state, action):  # :816-826
        acts = [action] if isinstance(action[0], int) else list(action)
        return self.time_cost + self.action_cost * sum(1 for a in acts if tuple(a) != (0, 0))

    @property
    def status(self):
        """0 solved, 1 nothing to do, 2 unreachable, 3 search budget exceeded, 4 unsupported"""
        return self._status

    def get_next_action(self, env, subtask, subtask_agent_names, other_agent_planners):
        """argmin_a Q(s, a) with uniform random tie-break (:987-1076); None when there is nothing to do"""
        self.set_settings(env, subtask, subtask_agent_names, other_agent_planners)
        if subtask is None or not self._q:
            return None
        actions = list(self._q.keys())
        return actions[argmin([self._q[a] for a in actions])]
