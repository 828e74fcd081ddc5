"""`BayesianDelegator` with the reference's call signatures, answered by the kernels.

Reference: delegation_planner/bayesian_delegator.py - ctor(agent_name, all_agent_names, model_type,
planner, none_action_prob) :19-47, should_reset_priors :54-79, set_priors :262-290, bayes_update
:1026-1072, select_subtask :1009-1017, get_other_agent_planners :375-429, `probs`
(SubtaskAllocDistribution, delegation_planner/utils.py:8-193), SubtaskAllocation namedtuple :14.

What runs where: the hypothesis space (add_subtasks :792-886, add_dc_subtasks :928-1000,
add_greedy_subtasks :892-923) is host combinatorics; "doable" (subtask_alloc_is_doable :98-156) is
gc_lower_bound; spatial priors (:296-369) and the likelihood Q's (:461-689) come from the exact
planner kernels through `E2E_BRTDP`; the Bayes update itself (:1045-1072) is gc_bd_posterior_f64.
Because the Q's are exact fixed points instead of the reference's partially refined BRTDP lower
bounds, posteriors differ from a reference RUN; given identical Q inputs they agree to 1e-12
(tests/test_bd_gpu.py).
"""
import copy
import random
from collections import namedtuple
from itertools import combinations, permutations

import numpy as np
import torch

from . import navigation_planner, planning, recipe_planner

SubtaskAllocation = namedtuple("SubtaskAllocation", "subtask subtask_agent_names")
UNREACHABLE_Q = 100.0


class SubtaskAllocDistribution:
    """Distribution over subtask allocations (delegation_planner/utils.py:8-193)."""

    def __init__(self, subtask_allocs):
        self.probs = {}
        if subtask_allocs:
            prior = 1.0 / len(subtask_allocs)
            for alloc in subtask_allocs:
                self.probs[tuple(alloc)] = prior

    def __str__(self):
        return "".join("%s: %s\n" % (a, p) for a, p in self.probs.items())

    def enumerate_subtask_allocs(self):
        return list(self.probs.keys())

    def get_list(self):
        return list(self.probs.items())

    def get(self, subtask_alloc):
        return self.probs[tuple(subtask_alloc)]

    def get_max(self):  # :37-42 uniform random among exact ties
        if not self.probs:
            return None
        best = max(self.probs.values())
        return random.choice([a for a, p in self.probs.items() if p == best])

    def get_max_bucketed(self):  # :50-158 - marginal over single allocations, best full allocation containing it
        subtasks, probs = [], []
        for alloc, p in self.probs.items():
            for t in alloc:
                if t in subtasks:
                    probs[subtasks.index(t)] += p
                else:
                    subtasks.append(t)
                    probs.append(p)
        if not subtasks:
            return None
        best = subtasks[int(np.argmax(probs))]
        cands = [(a, p) for a, p in self.probs.items() if best in a]
        return max(cands, key=lambda ap: ap[1])[0]

    def get_best_containing(self, subtask):  # :160-168 (returns the subtask itself, as the reference does)
        valid, valid_p = [], []
        for alloc, p in self.probs.items():
            if subtask in alloc:
                valid.append(subtask)
                valid_p.append(p)
        return valid[int(np.argmax(valid_p))]

    def set(self, subtask_alloc, value):
        self.probs[tuple(subtask_alloc)] = value

    def update(self, subtask_alloc, factor):
        self.probs[tuple(subtask_alloc)] *= factor

    def delete(self, subtask_alloc):
        self.probs.pop(tuple(subtask_alloc), None)

    def normalize(self):  # :186-193
        total = sum(self.probs.values())
        for a in self.probs:
            self.probs[a] = 1.0 / len(self.probs) if total == 0 else self.probs[a] / total
        return self.probs


# ---------------------------------------------------------------------------------------
# hypothesis space (pure combinatorics; compared with the reference's in tests/test_delegation.py)
# ---------------------------------------------------------------------------------------
def _others(remaining_agents, remaining_subtasks, base):
    """get_other_subtask_allocations :697-784"""
    if not remaining_agents:
        return [base]
    if not remaining_subtasks:
        return [base + [SubtaskAllocation(None, tuple(a))] for a in remaining_agents]
    out = [base + [SubtaskAllocation(t, tuple(remaining_agents))] for t in remaining_subtasks]
    if len(remaining_agents) > 1 and len(remaining_subtasks) > 1:
        for ts in permutations(remaining_subtasks, 2):
            out.append(base + [SubtaskAllocation(ts[0], (remaining_agents[0],)),
                               SubtaskAllocation(ts[1], (remaining_agents[1],))])
    return out


def hypothesis_space(model_type, agent_name, all_agent_names, incomplete_subtasks):
    """List of subtask allocations (each a list of SubtaskAllocation) before pruning."""
    names = list(all_agent_names)
    subtasks = list(incomplete_subtasks)
    if model_type == "greedy":  # :892-923
        if None not in subtasks:
            subtasks = subtasks + [None]
        return [[SubtaskAllocation(t, (agent_name,))] for t in subtasks]
    if model_type == "dc":  # :928-1000
        pool = subtasks + [None] * (len(names) - 1)
        return [[SubtaskAllocation(p[i], (names[i],)) for i in range(len(names))]
                for p in permutations(pool, len(names))]
    if len(names) == 1:  # :848-852
        return [[SubtaskAllocation(t, tuple(names))] for t in subtasks]
    allocs = []
    for first in combinations(names, 2):  # :856-885
        pool = subtasks + [None] * (len(names) - 1)
        rest_agents = sorted(set(names) - set(first))
        for t in pool:
            allocs += _others(rest_agents, list(set(pool) - {t}), [SubtaskAllocation(t, tuple(first))])
        if len(pool) > 1:
            for ts in permutations(pool, 2):
                allocs += _others(rest_agents, list(set(pool) - set(ts)),
                                  [SubtaskAllocation(ts[0], (first[0],)), SubtaskAllocation(ts[1], (first[1],))])
    return allocs


class BayesianDelegator:
    """Bayesian Delegation over subtask allocations."""

    def __init__(self, agent_name, all_agent_names, model_type, planner, none_action_prob):
        self.name = "Bayesian Delegator"
        self.agent_name = agent_name
        self.all_agent_names = all_agent_names
        self.probs = None
        self.model_type = model_type
        self.priors = "uniform" if model_type == "up" else "spatial"
        self.planner = planner
        self.none_action_prob = none_action_prob
        self.incomplete_subtasks = []

    # -- hypothesis space ------------------------------------------------------------------
    def get_subtask_alloc_probs(self):  # :81-90
        return SubtaskAllocDistribution(hypothesis_space(self.model_type, self.agent_name, self.all_agent_names,
                                                         self.incomplete_subtasks))

    def add_subtasks(self):  # :792-886
        return hypothesis_space("bd", self.agent_name, self.all_agent_names, self.incomplete_subtasks)

    def add_greedy_subtasks(self):  # :892-923
        return hypothesis_space("greedy", self.agent_name, self.all_agent_names, self.incomplete_subtasks)

    def add_dc_subtasks(self):  # :928-1000
        return hypothesis_space("dc", self.agent_name, self.all_agent_names, self.incomplete_subtasks)

    def get_other_subtask_allocations(self, remaining_agents, remaining_subtasks, base_subtask_alloc):  # :697-784
        return _others(list(remaining_agents), list(remaining_subtasks), list(base_subtask_alloc))

    def _doable_table(self, env, allocs):
        """{(subtask, agent names): doable} for every pair named by `allocs`, one gc_lower_bound call"""
        pairs = sorted({(t.subtask, tuple(t.subtask_agent_names)) for a in allocs for t in a if t.subtask is not None},
                       key=lambda p: (str(p[0]), p[1]))
        if not pairs:
            return {}
        pb = navigation_planner._plan_batch(env)
        words = np.array(navigation_planner.packed_words(env), dtype=np.uint32).view(np.int32)
        pb.state.copy_(torch.from_numpy(words).to(pb.device).view(1, 4))
        subtasks = sorted({p[0] for p in pairs}, key=str)
        pb.set_subtask_masks([recipe_planner.subtask_masks(s) for s in subtasks])
        triples = []
        for st, names in pairs:
            idx = sorted(int(nm.split("-")[1]) - 1 for nm in names)
            triples.append((subtasks.index(st), idx[0], idx[1] if len(idx) > 1 else None))
        lb = planning.lower_bound(pb, triples)[0].tolist()
        perimeter = env.world.perimeter
        return {p: lb[k] < perimeter for k, p in enumerate(pairs)}  # bd:156

    def subtask_alloc_is_doable(self, env, subtask, subtask_agent_names):  # :98-156
        if subtask is None:
            return True
        return self._doable_table(env, [[SubtaskAllocation(subtask, tuple(subtask_agent_names))]])[
            (subtask, tuple(subtask_agent_names))]

    def prune_subtask_allocs(self, observation, subtask_alloc_probs):  # :200-256
        allocs = subtask_alloc_probs.enumerate_subtask_allocs()
        doable = self._doable_table(observation, allocs)
        for alloc in allocs:
            for t in alloc:
                if t.subtask is not None and not doable[(t.subtask, tuple(t.subtask_agent_names))]:
                    subtask_alloc_probs.delete(alloc)
                    break
                if t.subtask is None and len(t.subtask_agent_names) > 1:
                    subtask_alloc_probs.delete(alloc)
                    break
            if all(t.subtask is None for t in alloc) and len(alloc) > 1:
                subtask_alloc_probs.delete(alloc)
        return subtask_alloc_probs

    def should_reset_priors(self, obs, incomplete_subtasks):  # :54-79
        if self.probs is None:
            return True
        self.incomplete_subtasks = incomplete_subtasks
        probs = self.prune_subtask_allocs(obs, self.get_subtask_alloc_probs())
        return len(self.probs.enumerate_subtask_allocs()) != len(probs.enumerate_subtask_allocs())

    # -- priors ------------------------------------------------------------------------------
    def get_lower_bound_for_subtask_alloc(self, obs, subtask, subtask_agent_names):  # :162-194
        if subtask is None:
            return 0
        self.planner.get_next_action(env=obs, subtask=subtask, subtask_agent_names=subtask_agent_names,
                                     other_agent_planners={})
        return self.planner.v_l[(self.planner.cur_state.get_repr(), subtask)]

    def get_spatial_priors(self, obs, some_probs):  # :296-369
        for alloc in some_probs.enumerate_subtask_allocs():
            total_weight = 0.0
            for t in alloc:
                if t.subtask is not None:
                    total_weight += 1.0 / float(self.get_lower_bound_for_subtask_alloc(obs, t.subtask, t.subtask_agent_names))
            some_probs.update(alloc, len(t) ** 2.0 * total_weight)  # len(t) is the namedtuple's 2 fields (:365-367)
        return some_probs

    def set_priors(self, obs, incomplete_subtasks, priors_type):  # :262-290
        self.incomplete_subtasks = incomplete_subtasks
        probs = self.prune_subtask_allocs(obs, self.get_subtask_alloc_probs())
        probs.normalize()
        self.probs = self.get_spatial_priors(obs, probs) if priors_type == "spatial" else probs
        self.ensure_at_least_one_subtask()
        self.probs.normalize()

    def ensure_at_least_one_subtask(self):  # :1019-1024
        if self.model_type in ("greedy", "dc") and not self.probs.probs:
            self.probs = SubtaskAllocDistribution([[SubtaskAllocation(None, (self.agent_name,))]])

    def select_subtask(self, agent_name):  # :1009-1017
        best = self.probs.get_max()
        if best is not None:
            for t in best:
                if agent_name in t.subtask_agent_names:
                    return t.subtask, t.subtask_agent_names
        return None, agent_name

    def get_other_agent_planners(self, obs, backup_subtask):  # :375-429
        planners = {}
        for other in self.all_agent_names:
            if other == self.agent_name:
                continue
            subtask, names = self.select_subtask(agent_name=other)
            if subtask is None:
                subtask, names = backup_subtask, tuple(sorted([other, self.agent_name]))
            planner = copy.copy(self.planner)
            planner.set_settings(env=obs, subtask=subtask, subtask_agent_names=names)
            planners[other] = planner
        return planners

    # -- likelihood + Bayes update -------------------------------------------------------------
    def likelihood_row(self, obs_tm1, actions_tm1, subtask, subtask_agent_names, no_level_1=False):
        """The softmax input of prob_nav_actions (:461-689) for one (subtask, agents): returns
        (qdiff list, index of the taken action).  beta is applied by the kernel."""
        names = tuple(subtask_agent_names)
        if subtask is None:  # :618-641
            assert len(names) != 2, "Two agents are doing None."
            me = next(a for a in obs_tm1.sim_agents if a.name == self.agent_name)
            k = len(_single_actions(obs_tm1, me)) - 1
            diffs = [self.none_action_prob] + [(1.0 - self.none_action_prob) / k] * k
            return diffs, (0 if tuple(actions_tm1[names[0]]) == (0, 0) else 1)
        action = tuple(tuple(actions_tm1[n]) for n in names)
        action = action[0] if len(names) == 1 else action
        others = {} if (no_level_1 or len(self.all_agent_names) == len(names)) else {"level-1": True}
        self.planner.set_settings(env=obs_tm1, subtask=subtask, subtask_agent_names=names, other_agent_planners=others)
        valid = self.planner.get_actions()
        assert action in valid, "valid_nav_actions: %s action: %s" % (valid, action)  # bd:672
        if len(names) == 2 and self.agent_name in names:  # :677-679 only joint actions matching the partner's move
            other = 1 - names.index(self.agent_name)
            valid = [a for a in valid if a[other] == action[other]]
        # an offered action from which the goal is out of reach has Q = +inf; cap it so that the
        # differences stay finite (the reference's heuristic values are always finite)
        q = [min(self.planner.Q(obs_tm1, a), UNREACHABLE_Q) for a in valid]
        old_q = q[valid.index(action)]
        return [old_q - qa for qa in q], valid.index(action)

    def prob_nav_actions(self, obs_tm1, actions_tm1, subtask, subtask_agent_names, beta, no_level_1):
        qd, idx = self.likelihood_row(obs_tm1, actions_tm1, subtask, subtask_agent_names, no_level_1)
        x = beta * np.asarray(qd, dtype=np.float64)
        e = np.exp(x - x.max())
        return float(e[idx] / e.sum())

    def bayes_update(self, obs_tm1, actions_tm1, beta):  # :1026-1072
        allocs = self.probs.enumerate_subtask_allocs()
        doable = self._doable_table(obs_tm1, allocs)
        for alloc in allocs:
            if any(t.subtask is not None and not doable[(t.subtask, tuple(t.subtask_agent_names))] for t in alloc):
                self.probs.delete(alloc)
        self.ensure_at_least_one_subtask()
        if self.model_type == "fb":
            return
        allocs = self.probs.enumerate_subtask_allocs()
        if not allocs:
            return
        # distinct likelihood rows, then ONE posterior update on the device (gc_bd_posterior_f64)
        rows, row_index, entries = [], {}, []
        for alloc in allocs:
            ent = []
            for t in alloc:
                if self.model_type == "greedy" and self.agent_name not in t.subtask_agent_names:
                    continue
                key = (t.subtask, tuple(t.subtask_agent_names))
                if key not in row_index:
                    qd, idx = self.likelihood_row(obs_tm1, actions_tm1, t.subtask, t.subtask_agent_names)
                    row_index[key] = len(rows)
                    rows.append((qd, idx, 1 if self.model_type == "greedy" else len(t.subtask_agent_names)))
                ent.append(row_index[key])
            entries.append(ent)
        dev = navigation_planner._plan_batch(obs_tm1).device
        H, P = len(allocs), len(rows)
        A, E = max(len(r[0]) for r in rows), max(1, max(len(e) for e in entries))
        probs = torch.tensor([[self.probs.get(a) for a in allocs]], dtype=torch.float64, device=dev)
        hyp = torch.full((1, H, E), 255, dtype=torch.uint8)
        for h, ent in enumerate(entries):
            hyp[0, h, :len(ent)] = torch.tensor(ent, dtype=torch.uint8)
        qd = torch.zeros((1, P, A), dtype=torch.float64)
        for p, (d, _, _) in enumerate(rows):
            qd[0, p, :len(d)] = torch.tensor(d, dtype=torch.float64)
        nv = torch.tensor([[len(r[0]) for r in rows]], dtype=torch.uint8)
        ai = torch.tensor([[r[1] for r in rows]], dtype=torch.uint8)
        w = torch.tensor([[r[2] for r in rows]], dtype=torch.uint8)
        planning.bd_posterior(probs, None, hyp.to(dev), w.to(dev), qd.to(dev), nv.to(dev), ai.to(dev), beta)
        for a, p in zip(allocs, probs[0].tolist()):
            self.probs.set(a, p)


def _single_actions(env, agent):
    """nav_utils.get_single_actions (navigation_planner/utils.py:55-90) on the host views."""
    from .utils.core import Delivery, Object, mergeable
    actions = []
    locs = [a.location for a in env.sim_agents]
    for t in [(0, 1), (0, -1), (-1, 0), (1, 0)]:
        new = env.world.inbounds((agent.location[0] + t[0], agent.location[1] + t[1]))
        if new in locs:
            continue
        gs = env.world.get_gridsquare_at(new)
        if not gs.collidable or isinstance(gs, Delivery):
            actions.append(t)
        elif gs.holding is None and agent.holding is not None:
            actions.append(t)
        elif isinstance(gs.holding, Object) and agent.holding is None:
            actions.append(t)
        elif isinstance(gs.holding, Object) and agent.holding is not None and mergeable(agent.holding, gs.holding):
            actions.append(t)
    actions.append((0, 0))
    return actions
