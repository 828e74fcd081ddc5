"""Device-resident Bayesian-Delegation loop: `main_loop` (main.py:85-117) for N kitchens at once.

Every env of a `KitchenBatch` carries its own two `RealAgent`s (utils/agent.py:28-368), each with its
own `BayesianDelegator` (delegation_planner/bayesian_delegator.py) - as tensors instead of Python
objects:

  reference object state                          here (N = envs, i = observer)
  ------------------------------------------------------------------------------------------
  RealAgent.incomplete_subtasks                   incomplete[N][i]   bit s = subtask s still open
  RealAgent.subtask / subtask_agent_names         cur_sub[N][i], cur_joint[N][i]
  BayesianDelegator.probs (dict alloc -> p)       probs[N][i][H] + alive[N][i][H] over the STATIC
                                                  hypothesis table of the level's full subtask list
  E2E_BRTDP.v_l / v_u dicts keyed by state repr   PlanCache: exact V*/Q rows keyed by planning state

The hypothesis table is `delegation_planner.hypothesis_space` (bd:792-1000) evaluated once on all
subtasks; what the reference regenerates per step from `incomplete_subtasks` and prunes with
`subtask_alloc_is_doable` (bd:98-156, 200-256) is the `alive` mask.  Per step and observer the order
of operations is RealAgent.select_action :82-104: update_subtasks (reset / set_priors / bayes_update,
:176-203) -> select_subtask (bd:1009-1017) -> plan (:218-281); after env.step, refresh_subtasks
(:151-171).  Hot numerics are the kernels (gc_lower_bound, gc_subtask_q/gc_joint_q through the
cache, gc_bd_posterior_f64, gc_env_step); the masks, gathers and arg-max/min around them are torch
tensor ops on the same device - nothing leaves HBM during an episode.
"""
import os

import numpy as np
import torch

from . import engine, planning, recipe_planner
from .delegation_planner import UNREACHABLE_Q, SubtaskAllocation, hypothesis_space

_T_MASK = ~0xFF000000  # clears t and done of w[0] inside the low int64 of a packed state
DENSE_MAX_H = 128       # hypothesis tables up to this size stay dense per env (gc_bd_posterior's limit)
_FORCE_LISTS = bool(os.environ.get("GC_BD_LISTS"))  # tests: run small tables through the list form too


class _StateView:
    """What planning.* needs from a KitchenBatch, over an arbitrary state tensor of the same level."""

    def __init__(self, batch):
        self._batch = batch
        self.num_agents, self.n_levels, self.level_id, self.device = batch.num_agents, batch.n_levels, None, batch.device
        self.state, self.num_envs = None, 0

    def _lv(self):
        return self._batch._lv()

    def _stream(self):
        return self._batch._stream()

    def on(self, state):
        self.state, self.num_envs = state, state.shape[0]
        return self


class PlanCache:
    """Exact planner answers memoised by planning state (packed state with t and done cleared):
    the batched form of the reference's `v_l`/`v_u` dicts keyed by `(state.get_repr(), subtask)`
    (e2e:216-352) - a state reached by many envs, or again on a later step, is solved once."""

    def __init__(self, batch, pairs):
        self.view = _StateView(batch)
        self.pairs = list(pairs)
        dev, P = batch.device, len(self.pairs)
        self.keys = torch.empty((0, 2), dtype=torch.int64, device=dev)
        self.row = torch.empty((0,), dtype=torch.int64, device=dev)
        self.v = torch.empty((0, P), dtype=torch.float32, device=dev)
        self.q = torch.empty((0, P, 25), dtype=torch.float32, device=dev)
        self.status = torch.empty((0, P), dtype=torch.uint8, device=dev)
        self.solved_states = 0
        self.lookups = 0

    @staticmethod
    def key_of(state):
        k = state.contiguous().view(torch.int64).clone()  # [N][2]: (w0 | w1 << 32, w2 | w3 << 32)
        k[:, 0] &= _T_MASK
        return k

    def lookup(self, state):
        """cache row of every env of `state` (int32[N][4]); unseen planning states are solved first.
        Rows are append-only, so indices returned earlier stay valid."""
        uk, inv = torch.unique(self.key_of(state), dim=0, return_inverse=True)
        self.lookups += state.shape[0]
        C = self.keys.shape[0]
        merged, pos = torch.unique(torch.cat([self.keys, uk]), dim=0, return_inverse=True)
        pos_old, pos_new = pos[:C], pos[C:]
        if merged.shape[0] > C:
            present = torch.zeros(merged.shape[0], dtype=torch.bool, device=state.device)
            present[pos_old] = True
            miss = ~present[pos_new]
            v, q, status = planning.subtask_q(self.view.on(uk[miss].contiguous().view(torch.int32)), self.pairs)
            n_miss = v.shape[0]
            self.solved_states += n_miss
            row = torch.empty(merged.shape[0], dtype=torch.int64, device=state.device)
            row[pos_old] = self.row
            row[pos_new[miss]] = C + torch.arange(n_miss, device=state.device)
            self.keys, self.row = merged, row  # sorted keys -> data row
            self.v, self.q, self.status = torch.cat([self.v, v]), torch.cat([self.q, q]), torch.cat([self.status, status])
        return self.row[pos_new][inv]


def alloc_key(alloc, subtasks):
    """canonical sort key of one subtask allocation (deterministic tie-breaking in tests)"""
    return str(sorted((len(subtasks) if t.subtask is None else subtasks.index(t.subtask),
                       tuple(sorted(int(nm.split("-")[1]) - 1 for nm in t.subtask_agent_names))) for t in alloc))


class _ObserverTables:
    """Static hypothesis / likelihood-row tables of one observer (agent index `me`, model type)."""

    def __init__(self, owner, me, model):
        S, names, dev = owner.S, owner.names, owner.device
        self.me, self.model = me, model
        self.spatial = model != "up"  # RealAgent.__init__ :52-55
        allocs = list(dict.fromkeys(tuple(a) for a in hypothesis_space(model, names[me], names, owner.subtasks)))
        if model == "dc":  # ensure_at_least_one_subtask (bd:1019-1024): used only when nothing else is alive
            allocs.append((SubtaskAllocation(None, (names[me],)),))
        H = len(allocs)
        E = max(len(a) for a in allocs)
        sub = np.full((H, E), -1, dtype=np.int64)      # subtask index, S = None, -1 = no entry
        lid = np.zeros((H, E), dtype=np.int64)          # index into owner.lpairs (doable table)
        pid0 = np.zeros((H, E), dtype=np.int64)         # level-0 planner pair (priors)
        hyp_pair = np.full((H, E), 255, dtype=np.uint8)
        static_ok = np.ones(H, dtype=bool)
        fallback = np.zeros(H, dtype=bool)
        sel_sub = np.full(H, S, dtype=np.int64)
        sel_joint = np.zeros(H, dtype=bool)
        sel_first = np.ones(H, dtype=bool)             # the observer is the first agent of its (sorted) pair
        plan_pid = np.zeros(H, dtype=np.int64)          # planner pair RealAgent.plan asks for (agent.py:244-263)
        # planning level of `plan` (e2e:360-406): greedy passes no other planners (level 0); otherwise level 1
        # as soon as somebody is outside the subtask's agent set
        plan_level = lambda ag: 0 if model == "greedy" else (1 if len(ag) < owner.NA else 0)
        rows, row_index = [], {}
        keys = []
        for h, alloc in enumerate(allocs):
            ents = []
            for e, t in enumerate(alloc):
                ag = tuple(sorted(int(nm.split("-")[1]) - 1 for nm in t.subtask_agent_names))
                s = S if t.subtask is None else owner.subtasks.index(t.subtask)
                sub[h, e] = s
                ents.append((s, ag))
                if s < S:
                    lid[h, e] = owner.lid[(s, ag)]
                    pid0[h, e] = owner.pid[(s, ag, 0)]
                elif len(ag) > 1:
                    static_ok[h] = False  # nobody does None together (bd:243-246)
                if me in ag:
                    sel_sub[h], sel_joint[h] = s, len(ag) > 1
                    sel_first[h] = ag[0] == me
                    if s < S:
                        plan_pid[h] = owner.pid[(s, ag, plan_level(ag))]
                if model == "greedy" and me not in ag:
                    continue
                if s == S and len(ag) > 1:
                    continue
                if (s, ag) not in row_index:
                    row_index[(s, ag)] = len(rows)
                    rows.append((s, ag))
                hyp_pair[h, e] = row_index[(s, ag)]
            if all(s == S for s, _ in ents) and len(ents) > 1:
                static_ok[h] = False  # at least one agent works (bd:249-253)
            keys.append(alloc_key(alloc, owner.subtasks))
        if model == "dc":
            fallback[H - 1] = True
            static_ok[H - 1] = False
        self.H, self.E, self.P = H, E, len(rows)
        self.allocs, self.keys, self.rows = allocs, keys, rows
        order = sorted(range(H), key=lambda h: keys[h])
        rank = np.empty(H, dtype=np.int64)
        rank[order] = np.arange(H)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.ent_sub, self.ent_lid, self.ent_pid0 = t(sub), t(lid), t(pid0)
        self.ent_has = t((sub >= 0) & (sub < S))
        has = (sub >= 0) & (sub < S)
        self.ent_lidx = t(np.where(has, lid, len(owner.lpairs)))
        self.ent_pidx = t(np.where(has, pid0, len(owner.cpairs)))
        self.ent_subc = t(np.clip(sub, 0, S - 1))
        self.hyp_pair = t(hyp_pair)
        self.static_ok, self.fallback, self.rank = t(static_ok), t(fallback), t(rank)
        self.sel_sub, self.sel_joint = t(sel_sub), t(sel_joint)
        self.sel_first, self.plan_pid = t(sel_first), t(plan_pid)
        # likelihood rows (bd:461-689): None rows (kind 0), single-agent rows (1), joint rows with the observer in
        # the pair (2: partner filter, five entries) or outside it (3: all 25 joint actions); the planning world
        # is level 1 as soon as somebody is outside the row's agent set (:650-654), else level 0
        kind = np.array([0 if s == S else (1 if len(ag) == 1 else (2 if me in ag else 3)) for s, ag in rows], dtype=np.int64)
        agent = np.array([ag[0] for s, ag in rows], dtype=np.int64)
        pid = np.array([0 if s == S else owner.pid[(s, ag, 1 if len(ag) < owner.NA else 0)] for s, ag in rows],
                       dtype=np.int64)
        w = np.array([1 if model == "greedy" else len(ag) for s, ag in rows], dtype=np.uint8)
        self.row_kind, self.row_agent, self.row_pid, self.pair_w = t(kind), t(agent), t(pid), t(w)
        self.row_kind_host, self.row_agent_host, self.row_pid_host = kind, agent, pid
        self.row_agent2_host = np.array([ag[1] if len(ag) > 1 else ag[0] for s, ag in rows], dtype=np.int64)
        self.A = 25 if bool((kind == 3).any()) else 5
        # list form (three / four agents): an env keeps the table rows that survive pruning as a list `rid`; row H is
        # the padding row (no entries, never alive)
        self.lists = H > DENSE_MAX_H or _FORCE_LISTS
        pad = lambda a, fill: t(np.concatenate([a, np.full((1,) + a.shape[1:], fill, dtype=a.dtype)]))
        self.ent_lidx_pad = pad(np.where(has, lid, len(owner.lpairs)), len(owner.lpairs))
        self.ent_pidx_pad = pad(np.where(has, pid0, len(owner.cpairs)), len(owner.cpairs))
        self.rank_pad = pad(rank, H)
        self.sel_sub_pad, self.sel_joint_pad = pad(sel_sub, S), pad(sel_joint, False)
        self.sel_first_pad, self.plan_pid_pad = pad(sel_first, True), pad(plan_pid, 0)
        self.hyp_pair_dev = self.hyp_pair.contiguous()
        self.pair_w_host = w


class BatchedDelegation:
    """N episodes of the reference's main loop with every agent a `RealAgent`, on one GPU."""

    def __init__(self, level, num_envs, models=("bd", "bd"), max_num_timesteps=100, beta=1.3, none_action_prob=0.5,
                 seed=1, deterministic=False, device=None):
        self.NA = len(models)
        if not 1 <= self.NA <= 4:
            raise ValueError("1 to 4 agents")
        for m in models:
            if m not in ("bd", "up", "fb", "dc", "greedy"):
                raise ValueError("unknown model type %r" % (m,))
        self.models, self.beta, self.none_action_prob = tuple(models), float(beta), float(none_action_prob)
        self.deterministic = bool(deterministic)
        self._level, self._max_t = level, max_num_timesteps
        self.kb = engine.KitchenBatch(level, self.NA, num_envs, max_num_timesteps, device=device, track_collisions=True)
        self.device, self.N = self.kb.device, self.kb.num_envs
        self.names = ["agent-%d" % (i + 1) for i in range(self.NA)]
        self.subtasks = list(self.kb.subtasks[0])
        self.S = S = len(self.subtasks)
        lv = self.kb.levels[0]
        self.perimeter = float(2 * (lv.width + lv.height))  # world.perimeter, env:198
        import itertools
        agsets = [(i,) for i in range(self.NA)] + list(itertools.combinations(range(self.NA), 2))
        self.lpairs, self.lid, self.cpairs, self.pid = [], {}, [], {}
        for s in range(S):
            for ag in agsets:
                self.lid[(s, ag)] = len(self.lpairs)
                self.lpairs.append((s, ag[0], ag[1] if len(ag) > 1 else None))
                # level 0 (priors, greedy planning; everybody inside the set) and, when somebody is outside the
                # agent set, level 1 (likelihoods and non-greedy planning)
                for lvl in ((0, 1) if len(ag) < self.NA else (0,)):
                    self.pid[(s, ag, lvl)] = len(self.cpairs)
                    self.cpairs.append((s, ag[0], ag[1] if len(ag) > 1 else None, bool(lvl)))
        self.lpair_sub = torch.tensor([p[0] for p in self.lpairs], dtype=torch.int64, device=self.device)
        self.cache = PlanCache(self.kb, self.cpairs)
        self.tables = [_ObserverTables(self, i, models[i]) for i in range(self.NA)]
        # the hypothesis space is tabulated over the level's full subtask list (H rows per observer); an env keeps a
        # dense probability row over it for two agents (H <= 100) and a list of the surviving rows beyond that
        # (H = 2 457 for three, 39 906 for four bd agents on a salad level, of which a few hundred survive pruning)
        masks = [recipe_planner.subtask_masks(s) for s in self.subtasks]
        dev = self.device
        self.goal_mask = torch.tensor([m[3] for m in masks], dtype=torch.int64, device=dev)
        self.is_deliver = torch.tensor([m[0] == recipe_planner.ST_DELIVER for m in masks], dtype=torch.bool, device=dev)
        cell_type = torch.tensor(list(lv.cell_type), dtype=torch.int64, device=dev)
        self.cell_type = cell_type
        self.is_delivery = cell_type == 3
        done_food = lambda m: ((m & 7) & ~(m >> 4)) == 0
        mg = [[(not (a & b & 8)) and done_food(a | b) for b in range(128)] for a in range(128)]  # core.mergeable
        self.mergeable = torch.tensor(mg, dtype=torch.bool, device=dev)
        self.delta = torch.tensor([8, -8, -1, 1], dtype=torch.int64, device=dev)
        self._four = torch.arange(4, dtype=torch.uint8, device=dev)[None, None, :]
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(int(seed))
        self.reset()

    # -- state -------------------------------------------------------------------------------
    def reset(self):
        self.wkb, self.ids = self.kb, None  # working batch: the envs still running (== kb until run() compacts)
        self.N = self.kb.num_envs
        N, NA, dev = self.N, self.NA, self.device
        self.kb.reset()
        self.t = 0
        self.incomplete = torch.full((N, NA), (1 << self.S) - 1, dtype=torch.int64, device=dev)
        self.cur_sub = torch.full((N, NA), self.S, dtype=torch.int64, device=dev)
        self.cur_joint = torch.zeros((N, NA), dtype=torch.bool, device=dev)
        self.has_probs = torch.zeros((N, NA), dtype=torch.bool, device=dev)
        self.probs = [torch.zeros((N, 1 if T.lists else T.H), dtype=torch.float64, device=dev) for T in self.tables]
        self.alive = [torch.zeros((N, 1 if T.lists else T.H), dtype=torch.bool, device=dev) for T in self.tables]
        self.rid = [torch.full((N, 1), T.H, dtype=torch.int64, device=dev) if T.lists else None for T in self.tables]
        self.executed = torch.full((N, NA), 4, dtype=torch.uint8, device=dev)
        self.prev = None
        self.posterior_updates = 0
        self.agent_steps = 0  # counted by run(): live envs x agents per loop step

    def _slots(self, state):
        """(mask, cell, holder) int64[N][6] of the six object slots"""
        w = state.to(torch.int64) & 0xFFFFFFFF  # byte planes: include/gymcook.h
        place = torch.stack([(w[:, 1] >> (8 * k)) & 0xFF for k in range(4)] +
                            [(w[:, 3] >> (8 * k)) & 0xFF for k in range(2)], dim=1)
        mask = torch.stack([(w[:, 2] >> (8 * k)) & 0xFF for k in range(4)] +
                           [(w[:, 3] >> (8 * (k + 2))) & 0xFF for k in range(2)], dim=1)
        held = place >= 0x40
        return mask, torch.where(held, torch.zeros_like(place), place), torch.where(held, place & 7, torch.zeros_like(place))

    def _agent_cells(self, state):
        w0 = state[:, 0].to(torch.int64) & 0xFFFFFFFF
        return torch.stack([(w0 >> (6 * i)) & 63 for i in range(self.NA)], dim=1)

    def single_actions(self, state):
        """nav_utils.get_single_actions (navigation_planner/utils.py:55-90) on the real env:
        bool[N][NA][4] for the four moves (stay is always offered)."""
        N = state.shape[0]
        mask, cell, holder = self._slots(state)
        cells = self._agent_cells(state)
        lying = holder == 0
        occ = torch.zeros((N, 64), dtype=torch.int64, device=state.device)
        occ.scatter_add_(1, cell, mask * lying)
        out = torch.zeros((N, self.NA, 4), dtype=torch.bool, device=state.device)
        for i in range(self.NA):
            hold = (mask * (holder == i + 1)).sum(1)
            for a in range(4):
                tgt = (cells[:, i] + self.delta[a]) & 63
                kind = self.cell_type[tgt]
                blocked = torch.zeros(N, dtype=torch.bool, device=state.device)
                for j in range(self.NA):
                    blocked |= cells[:, j] == tgt  # `if new_loc in agent_locs: continue`
                mT = occ.gather(1, tgt[:, None])[:, 0]
                counter = ((mT == 0) & (hold != 0)) | ((mT != 0) & (hold == 0)) | (
                    (mT != 0) & (hold != 0) & self.mergeable[hold, mT])
                out[:, i, a] = ~blocked & ((kind == 0) | (kind == 3) | counter)
        return out

    def goal_count(self, state, sub):
        """RealAgent.def_subtask_completion (agent.py:286-368): objects equal to the subtask's goal
        (for Deliver: lying on a delivery square), per env for that env's subtask `sub` (< S)"""
        mask, cell, holder = self._slots(state)
        goal = self.goal_mask[sub][:, None]
        match = (holder != 7) & (mask == goal)
        delivered = match & (holder == 0) & self.is_delivery[cell]
        return torch.where(self.is_deliver[sub], delivered.sum(1), match.sum(1))

    # -- one observer --------------------------------------------------------------------------
    def _entries_ok(self, T, doable, inc):
        """hypothesis h survives iff each of its entries is None / empty or names a doable (and, with
        `inc`, still incomplete) subtask: evaluated per (subtask, agent set) pair first ([N][3S], small),
        then gathered per hypothesis entry"""
        ok = doable
        if inc is not None:
            ok = ok & (((inc[:, None] >> self.lpair_sub[None, :]) & 1) != 0)
        ok = torch.cat([ok, torch.ones((ok.shape[0], 1), dtype=torch.bool, device=ok.device)], dim=1)
        out = ok[:, T.ent_lidx[:, 0]]
        for e in range(1, T.E):
            out = out & ok[:, T.ent_lidx[:, e]]
        return out

    def _pick(self, cand, rank=None):
        """index of one True per row of `cand`: uniformly random, or lowest rank when deterministic"""
        if self.deterministic:
            r = rank if rank is not None else torch.arange(cand.shape[1], device=cand.device)
            return torch.where(cand, r.expand_as(cand), 1 << 40).argmin(1)
        noise = torch.rand(cand.shape, device=cand.device, generator=self.gen) + 1e-6
        return (noise * cand).argmax(1)

    def _priors(self, T, alive, ci):
        cnt = alive.sum(1, keepdim=True).clamp(min=1)
        p = alive.double() / cnt
        if T.spatial:  # get_spatial_priors bd:296-369: 4 * sum_t 1 / v_l(t)
            inv_v = 1.0 / self.cache.v[ci].double().clamp(min=1e-9)
            inv_v = torch.cat([inv_v, torch.zeros((inv_v.shape[0], 1), dtype=torch.float64, device=inv_v.device)], dim=1)
            w = inv_v[:, T.ent_pidx[:, 0]]
            for e in range(1, T.E):
                w = w + inv_v[:, T.ent_pidx[:, e]]
            p = p * (w * 4.0)
        tot = p.sum(1, keepdim=True)
        return torch.where(tot == 0, alive.double() / cnt, p / tot.clamp(min=1e-300))

    def _likelihood_rows_torch(self, T):
        """torch restatement of gc_bd_likelihood_rows (kept as the kernel's cross-check in the tests)"""
        N, dev, me = self.N, self.device, T.me
        pv = self.prev
        ex = self.executed.long()
        taken = ex[:, T.row_agent]  # [N][P] action of the row's (first) agent
        ai = torch.arange(5, device=dev)[None, None, :]
        if self.NA == 2:  # joint rows: only joint actions matching the partner's move (bd:677-679)
            partner = ex[:, 1 - me][:, None, None]
            jidx = (ai * 5 + partner) if me == 0 else (partner * 5 + ai)
            is_joint = (T.row_kind == 2)[None, :, None]
            idx5 = torch.where(is_joint, jidx, ai)
            taken = torch.where(T.row_kind[None, :] == 2, ex[:, me][:, None].expand(N, T.P), taken)
        else:
            idx5 = ai.expand(N, T.P, 5)
        q5 = self.cache.q[pv["ci"][:, None, None], T.row_pid[None, :, None], idx5]  # [N][P][5]
        onehot = torch.arange(5, device=dev)[None, None, :] == taken[:, :, None]
        valid = ~torch.isnan(q5) | onehot
        qc = torch.nan_to_num(q5, nan=UNREACHABLE_Q, posinf=UNREACHABLE_Q).clamp(max=UNREACHABLE_Q).double()
        old = (qc * onehot).sum(-1, keepdim=True)
        qdiff = old - qc
        # None rows (bd:618-641): [p_none, (1 - p_none) / k, ...] with k = the OBSERVER's move count
        k = pv["offered"][:, me].sum(-1)  # [N]
        none_row = torch.cat([torch.full((N, 1), self.none_action_prob, dtype=torch.float64, device=dev),
                              ((1.0 - self.none_action_prob) / k.clamp(min=1).double())[:, None].expand(N, 4)], dim=1)
        is_none = (T.row_kind == 0)[None, :, None]
        none_valid = torch.arange(5, device=dev)[None, :] <= k[:, None]
        valid = torch.where(is_none, none_valid[:, None, :], valid)
        qdiff = torch.where(is_none, none_row[:, None, :], qdiff)
        # compact the valid actions to the front, in order (the kernel reads the first n_valid entries)
        before = torch.ones((5, 5), dtype=torch.bool, device=dev).tril(-1)  # before[k][j] = j < k
        n_valid = valid.sum(-1)
        rank_valid = (valid[:, :, None, :] & before).sum(-1)
        rank_invalid = (~valid[:, :, None, :] & before).sum(-1)
        dest = torch.where(valid, rank_valid, n_valid[:, :, None] + rank_invalid)
        qdiff = torch.zeros_like(qdiff).scatter_(2, dest, qdiff)
        act_idx = rank_valid.gather(2, taken[:, :, None])[:, :, 0]
        none_idx = torch.where(taken == 4, 0, 1).clamp(max=(n_valid - 1).clamp(min=0))
        act_idx = torch.where(T.row_kind[None, :] == 0, none_idx, act_idx)
        return qdiff.contiguous(), n_valid.to(torch.uint8).contiguous(), act_idx.to(torch.uint8).contiguous()

    def _likelihood_rows(self, T):
        """gc_bd_likelihood_rows on obs_tm1 = self.prev, actions_tm1 = self.executed"""
        pv = self.prev
        n_moves = pv["offered"][:, T.me].sum(-1).to(torch.uint8)
        return planning.bd_likelihood_rows(self.cache.q, pv["ci"].contiguous(), T.row_pid_host, T.row_kind_host,
                                           T.row_agent_host, T.row_agent2_host, self.executed, n_moves, T.me,
                                           self.none_action_prob, UNREACHABLE_Q)

    def _bayes_update(self, T, probs, alive):
        """bayes_update bd:1045-1072 on obs_tm1 = self.prev, actions_tm1 = self.executed"""
        N = self.N
        qdiff, n_valid, act_idx = self._likelihood_rows(T)
        out = (probs * alive).contiguous()
        if getattr(T, "_hyp_pair_n", None) is None or T._hyp_pair_n.shape[0] != N:  # static: expand once
            T._hyp_pair_n = T.hyp_pair[None].expand(N, T.H, T.E).contiguous()
            T._pair_w_n = T.pair_w[None].expand(N, T.P).contiguous()
        planning.bd_posterior(out, alive.to(torch.uint8).contiguous(), T._hyp_pair_n, T._pair_w_n, qdiff,
                              n_valid, act_idx, self.beta)
        self.posterior_updates += N
        return out

    def _select_action(self, i, ci, doable, offered):
        T, S = self.tables[i], self.S
        if T.lists:
            return self._select_action_lists(i, ci, doable, offered)
        inc, cur = self.incomplete[:, i], self.cur_sub[:, i]
        alive_new = self._entries_ok(T, doable, inc) & T.static_ok
        alive_cur, has = self.alive[i], self.has_probs[:, i]
        # update_subtasks :176-203
        stale = (cur < S) & (((inc >> cur.clamp(max=S - 1)) & 1) == 0)
        reset = stale | ~has | (alive_new.sum(1) != alive_cur.sum(1))  # should_reset_priors bd:54-79
        do_prior = reset | (cur >= S)
        if T.fallback.any():  # dc: an empty distribution becomes {None: me} (bd:1019-1024)
            alive_new = alive_new | (T.fallback[None, :] & ~alive_new.any(1, keepdim=True))
        prior = self._priors(T, alive_new, ci)
        if self.prev is not None and not bool(do_prior.all()):
            alive_upd = alive_cur & self._entries_ok(T, self.prev["doable"], None)
            if T.fallback.any():
                alive_upd = alive_upd | (T.fallback[None, :] & ~alive_upd.any(1, keepdim=True))
            if T.model == "fb":
                upd = self.probs[i] * alive_upd
            else:
                upd = self._bayes_update(T, self.probs[i], alive_upd)
            probs = torch.where(do_prior[:, None], prior, upd)
            alive = torch.where(do_prior[:, None], alive_new, alive_upd)
        else:
            probs, alive = prior, alive_new
        self.probs[i], self.alive[i] = probs, alive
        self.has_probs[:, i] = True
        # select_subtask bd:1009-1017 (get_max: uniform among exact ties, dutils:37-42)
        masked = torch.where(alive, probs, -1.0)
        best_p = masked.max(1, keepdim=True).values
        tol = 1e-12 if self.deterministic else 0.0
        best = self._pick(alive & (masked >= best_p - tol), T.rank)
        return self._plan(i, best, ~alive.any(1), ci, offered)

    # -- the list form: three and four agents ----------------------------------------------------
    def _alive_rows(self, T, ok):
        """Rows of T's hypothesis table whose every entry names an `ok` (subtask, agent set) pair, per env, as a
        front-packed list: (rows int64[N][Wn] padded with H, count int64[N]).  The surviving set depends on the env
        only through its `ok` bits, so it is evaluated once per DISTINCT bit pattern of the batch (bd:792-886
        regenerates and bd:200-256 prunes the space per agent and step)."""
        N, L = ok.shape
        dev = ok.device
        words = []
        for lo in range(0, L, 62):
            bits = ok[:, lo:lo + 62].long()
            words.append((bits << torch.arange(bits.shape[1], device=dev)[None, :]).sum(1))
        uk, inv = torch.unique(torch.stack(words, dim=1), dim=0, return_inverse=True)
        U = uk.shape[0]
        first = torch.zeros(U, dtype=torch.int64, device=dev)
        first.scatter_(0, inv, torch.arange(N, device=dev))  # any env of each pattern
        ok_u = torch.cat([ok[first], torch.ones((U, 1), dtype=torch.bool, device=dev)], dim=1)  # [U][L + 1]
        parts, counts = [], []
        chunk = max(1, (1 << 26) // T.H)
        for lo in range(0, U, chunk):
            o = ok_u[lo:lo + chunk]
            alive_u = T.static_ok[None, :] & o[:, T.ent_lidx[:, 0]]
            for e in range(1, T.E):
                alive_u &= o[:, T.ent_lidx[:, e]]
            k_u = alive_u.sum(1)
            nz = alive_u.nonzero()  # row-major: grouped by pattern, rows ascending
            parts.append((nz, k_u))
            counts.append(k_u)
        k_u = torch.cat(counts)
        Wn = max(int(k_u.max()), 1)
        rows_u = torch.full((U, Wn), T.H, dtype=torch.int64, device=dev)
        lo = 0
        for nz, k in parts:
            off = torch.cumsum(k, 0) - k
            j = torch.arange(nz.shape[0], device=dev) - off[nz[:, 0]]
            rows_u[lo + nz[:, 0], j] = nz[:, 1]
            lo += k.shape[0]
        return rows_u[inv], k_u[inv]

    @staticmethod
    def _widen(x, W, fill):
        if x.shape[1] >= W:
            return x
        return torch.cat([x, torch.full((x.shape[0], W - x.shape[1]), fill, dtype=x.dtype, device=x.device)], dim=1)

    def _select_action_lists(self, i, ci, doable, offered):
        """_select_action with the observer's distribution kept as a list of table rows per env"""
        T, S, N, dev = self.tables[i], self.S, self.N, self.device
        H = T.H
        inc, cur = self.incomplete[:, i], self.cur_sub[:, i]
        ok = doable & (((inc[:, None] >> self.lpair_sub[None, :]) & 1) != 0)
        rows_new, k_new = self._alive_rows(T, ok)
        alive_cur, rid_cur, has = self.alive[i], self.rid[i], self.has_probs[:, i]
        stale = (cur < S) & (((inc >> cur.clamp(max=S - 1)) & 1) == 0)
        reset = stale | ~has | (k_new != alive_cur.sum(1))  # should_reset_priors bd:54-79
        do_prior = reset | (cur >= S)
        has_fallback = bool(T.fallback.any())
        if has_fallback:  # dc: an empty distribution becomes {None: me} (bd:1019-1024); the fallback is row H - 1
            empty = k_new == 0
            rows_new[:, 0] = torch.where(empty, H - 1, rows_new[:, 0])
            k_new = torch.where(empty, 1, k_new)
        alive_new = torch.arange(rows_new.shape[1], device=dev)[None, :] < k_new[:, None]
        # priors (set_priors bd:262-290, get_spatial_priors :296-369)
        cnt = k_new.clamp(min=1)[:, None]
        prior = alive_new.double() / cnt
        if T.spatial:
            inv_v = 1.0 / self.cache.v[ci].double().clamp(min=1e-9)
            inv_v = torch.cat([inv_v, torch.zeros((N, 1), dtype=torch.float64, device=dev)], dim=1)
            w = inv_v.gather(1, T.ent_pidx_pad[:, 0][rows_new])
            for e in range(1, T.E):
                w = w + inv_v.gather(1, T.ent_pidx_pad[:, e][rows_new])
            p = prior * (w * 4.0)
            tot = p.sum(1, keepdim=True)
            prior = torch.where(tot == 0, prior, p / tot.clamp(min=1e-300))
        if self.prev is not None and not bool(do_prior.all()):
            okprev = torch.cat([self.prev["doable"], torch.ones((N, 1), dtype=torch.bool, device=dev)], dim=1)
            alive_upd = alive_cur.clone()
            for e in range(T.E):
                alive_upd &= okprev.gather(1, T.ent_lidx_pad[:, e][rid_cur])
            probs_cur = self.probs[i]
            if has_fallback:
                empty = ~alive_upd.any(1)
                in_list = ((rid_cur == H - 1) & empty[:, None])
                slot = torch.where(in_list.any(1), in_list.int().argmax(1), 0)  # its own slot if listed, else slot 0
                at = torch.arange(rid_cur.shape[1], device=dev)[None, :] == slot[:, None]
                put = empty[:, None] & at
                probs_cur = torch.where(put & ~in_list, 0.0, probs_cur)
                rid_cur = torch.where(put, H - 1, rid_cur)
                alive_upd = alive_upd | put
            if T.model == "fb":
                upd = probs_cur * alive_upd
            else:
                pv = self.prev
                upd = probs_cur.contiguous().clone()
                planning.bd_update_lists(upd, alive_upd.contiguous().view(torch.uint8), rid_cur.contiguous(), T.hyp_pair_dev,
                                         T.pair_w_host, self.cache.q, pv["ci"].contiguous(), T.row_pid_host,
                                         T.row_kind_host, T.row_agent_host, T.row_agent2_host, self.executed,
                                         pv["offered"][:, T.me].sum(-1).to(torch.uint8), T.me, self.none_action_prob,
                                         self.beta, UNREACHABLE_Q)
                self.posterior_updates += N
            W = max(rows_new.shape[1], rid_cur.shape[1])
            sel = do_prior[:, None]
            probs = torch.where(sel, self._widen(prior, W, 0.0), self._widen(upd, W, 0.0))
            alive = torch.where(sel, self._widen(alive_new, W, False), self._widen(alive_upd, W, False))
            rid = torch.where(sel, self._widen(rows_new, W, H), self._widen(rid_cur, W, H))
            used = alive.any(0).nonzero()
            Wt = int(used.max()) + 1 if used.numel() else 1  # drop the columns nobody uses any more
            probs, alive, rid = probs[:, :Wt].contiguous(), alive[:, :Wt].contiguous(), rid[:, :Wt].contiguous()
        else:
            probs, alive, rid = prior, alive_new, rows_new
        self.probs[i], self.alive[i], self.rid[i] = probs, alive, rid
        self.has_probs[:, i] = True
        masked = torch.where(alive, probs, -1.0)
        best_p = masked.max(1, keepdim=True).values
        tol = 1e-12 if self.deterministic else 0.0
        best = self._pick(alive & (masked >= best_p - tol), T.rank_pad[rid] if self.deterministic else None)
        nothing = ~alive.any(1)
        hbest = torch.where(nothing, H, rid.gather(1, best[:, None])[:, 0])
        return self._plan(i, hbest, nothing, ci, offered)

    def _plan(self, i, best, nothing, ci, offered):
        """select_subtask's result (table row `best`, H = the padding row) -> RealAgent.plan :218-281"""
        T, S, N, dev = self.tables[i], self.S, self.N, self.device
        new_sub = torch.where(nothing, S, T.sel_sub_pad[best])
        new_joint = T.sel_joint_pad[best] & ~nothing
        pid = T.plan_pid_pad[best]
        q = self.cache.q[ci, pid]  # [N][25]
        q = torch.where(new_joint[:, None] | (torch.arange(25, device=dev)[None, :] < 5), q, float("nan"))
        valid = ~torch.isnan(q)
        qv = torch.where(valid, q.clamp(max=1e30), float("inf"))
        a = self._pick(valid & (qv == qv.min(1, keepdim=True).values))  # argmin, random ties (e2e:27-30)
        own = torch.where(new_joint, torch.where(T.sel_first_pad[best], a // 5, a % 5), a)  # agent.py:270-272
        own = torch.where(valid.any(1), own, 4)
        # doing nothing: stay with none_action_prob, else a uniformly random offered move (:235-243)
        off = offered[:, i]
        if self.deterministic:
            a_none = torch.full((N,), 4, dtype=torch.int64, device=dev)
        else:
            u = torch.rand(N, device=dev, generator=self.gen)
            a_none = torch.where((u < self.none_action_prob) | ~off.any(1), 4, self._pick(off))
        self.cur_sub[:, i], self.cur_joint[:, i] = new_sub, new_joint
        return torch.where(new_sub >= S, a_none, own).to(torch.uint8)

    # -- the loop -----------------------------------------------------------------------------
    def step(self):
        """One pass of main_loop's body for every env of the working batch; returns its reward/done bytes."""
        kb = self.wkb
        state = kb.state
        ci = self.cache.lookup(state)
        doable = planning.lower_bound(kb, self.lpairs) < self.perimeter  # bd:156
        bits = planning.offered_actions(kb)  # get_single_actions on the real env; `single_actions` is its torch twin
        offered = ((bits[:, :, None] >> self._four) & 1).bool()
        actions = torch.stack([self._select_action(i, ci, doable, offered) for i in range(self.NA)], dim=1)
        self.prev = dict(state=state.clone(), ci=ci, doable=doable, offered=offered)
        self.last_actions = actions.contiguous()
        kb.step(self.last_actions, executed_out=self.executed)
        # refresh_subtasks :151-171 (gc_subtasks_completed; `goal_count` is its torch twin)
        done = planning.subtasks_completed(kb, self.prev["state"], self.cur_sub.to(torch.uint8).contiguous())
        self.incomplete &= ~(done.long() << self.cur_sub.clamp(max=self.S - 1))
        self.t += 1
        return kb.reward_done

    def _write_back(self):
        """final states of the working batch into the full batch"""
        if self.ids is not None:
            self.kb.state[self.ids] = self.wkb.state
            self.kb.reward_done[self.ids] = self.wkb.reward_done
            self.kb.collisions[self.ids] = self.wkb.collisions

    def _compact(self):
        """Drop the finished envs from the working batch: a done env only repeats its last state, but
        it would still pay for the delegation arithmetic of every later step."""
        self._write_back()
        keep = ~self.wkb.done
        live = keep.nonzero()[:, 0]
        ids = live if self.ids is None else self.ids[live]
        nkb = engine.KitchenBatch(self._level, self.NA, int(live.numel()), self._max_t, device=self.device,
                                  track_collisions=True)
        nkb.state.copy_(self.wkb.state[live])
        nkb.reward_done.copy_(self.wkb.reward_done[live])
        nkb.collisions.copy_(self.wkb.collisions[live])
        self.wkb, self.ids, self.N = nkb, ids, nkb.num_envs
        for name in ("incomplete", "cur_sub", "cur_joint", "has_probs", "executed"):
            setattr(self, name, getattr(self, name)[live].contiguous())
        self.probs = [p[live].contiguous() for p in self.probs]
        self.alive = [a[live].contiguous() for a in self.alive]
        self.rid = [None if r is None else r[live].contiguous() for r in self.rid]
        if self.prev is not None:
            self.prev = {k: v[live].contiguous() for k, v in self.prev.items()}

    def run(self, max_steps=None, compact=True):
        """Steps until every env is done (env.done(), main.py:99); returns the number of steps.  With
        `compact`, finished envs leave the working batch once they are the majority; `kb` holds
        every env's final state when this returns."""
        limit = max_steps if max_steps is not None else self.kb.max_num_timesteps + 1
        steps = 0
        done = int((self.wkb.reward_done & 1).sum()) if getattr(self, "wkb", None) is not None else 0
        while steps < limit:
            self.agent_steps += (self.N - done) * self.NA  # envs still running take this step
            rd = self.step()
            steps += 1
            done = int((rd & 1).sum())
            if done == self.N:
                break
            if compact and self.N >= 4096 and 2 * done >= self.N:
                self._compact()
                done = 0  # the working batch now holds running envs only
        self._write_back()
        return steps


# -- cfg-5 of the north star: four agents, mixed model types, all nine levels ---------------------------
MIXED_MODELS = ("bd", "up", "dc", "fb", "greedy")
MIXED_LEVELS = tuple("%s-divider_%s" % (d, r) for r in ("tomato", "tl", "salad") for d in ("open", "partial", "full"))


def run_mixed(envs_per_level, n_agents=4, levels=MIXED_LEVELS, shard=0, horizon=100, device=None, seed=1, max_steps=None):
    """The reference's experiment grid (run_experiments.sh: every level x every model type) as batches: for level
    k of `levels`, `envs_per_level` episodes of `n_agents` RealAgents whose model types are MIXED_MODELS rotated by
    k + shard (so every type meets every level and every seat).  Episodes are independent: `shard` (a rank) only
    changes seeds and the rotation.  Returns the totals and one record per level; wall-clock seconds are measured
    with a device synchronisation on both sides."""
    import time
    total = dict(envs=0, agent_steps=0, posterior_updates=0, delivered=0, seconds=0.0, planning_states_solved=0,
                 planner_lookups=0, completed_subtasks=0, per_level=[])
    for k, level in enumerate(levels):
        models = tuple(MIXED_MODELS[(k + shard + j) % len(MIXED_MODELS)] for j in range(n_agents))
        loop = BatchedDelegation(level, envs_per_level, models, max_num_timesteps=horizon,
                                 seed=seed + 1000 * shard + k, device=device)
        torch.cuda.synchronize(loop.device)
        t0 = time.perf_counter()
        steps = loop.run(max_steps=max_steps)
        torch.cuda.synchronize(loop.device)
        dt = time.perf_counter() - t0
        st = loop.kb.stats().cpu().tolist()
        rec = dict(level=level, models="/".join(models), hypotheses=[T.H for T in loop.tables], loop_steps=steps,
                   seconds=dt, agent_steps=loop.agent_steps, posterior_updates=loop.posterior_updates, delivered=st[1],
                   completed_subtasks=st[133], planning_states_solved=loop.cache.solved_states,
                   planner_lookups=loop.cache.lookups)
        total["per_level"].append(rec)
        total["envs"] += envs_per_level
        for key in ("agent_steps", "posterior_updates", "delivered", "seconds", "planning_states_solved", "planner_lookups",
                    "completed_subtasks"):
            total[key] += rec[key]
        del loop
    return total
