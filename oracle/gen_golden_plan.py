"""Golden fixtures for paths B and C, produced by running the unmodified reference.
Invoked through oracle/gen_golden.py (`lb`, `brtdp`, `bd`).  Build container only.

  lb     env.get_lower_bound_for_subtask_given_objs (env:594-664) on sampled states, for every
         (subtask, agent set)                                  -> tests/golden/lower_bounds.npz
  brtdp  E2E_BRTDP.get_next_action(..., other_agent_planners={}) (e2e_brtdp.py:987-1076): v_l, v_u at
         cur_state and Q_l per action                          -> tests/golden/brtdp_values.npz
  bd     BayesianDelegator.bayes_update (bd:1026-1072): prior, per-(alloc, t) softmax inputs, taken
         action index, weights -> posterior                    -> tests/golden/bd_posteriors.npz
"""
import copy
import itertools
import os
import random
import sys
import time
from multiprocessing import Pool

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import ref_harness as H  # noqa: E402
from gen_golden import GOLDEN, LEVELS, DELTA, walker_actions  # noqa: E402

ST_CHOP, ST_MERGE, ST_DELIVER = 1, 2, 3


def pack_env(env):
    """reference env -> packed uint32[4] (include/gymcook.h) with slots in canonical key order."""
    t, agents, items = H.canonical(env)
    w = [0, 0, 0, 0]
    for i, (x, y, _) in enumerate(agents):
        w[0] |= (y * 8 + x) << (6 * i)
    w[0] |= (t & 127) << 24
    slots = []
    used = [False] * len(agents)
    for (m, x, y, held) in sorted(items, key=lambda it: (it[0] << 7) | ((it[2] * 8 + it[1]) << 1) | it[3]):
        if held:
            i = next(i for i, (ax, ay, hm) in enumerate(agents) if (ax, ay) == (x, y) and hm == m and not used[i])
            used[i] = True
            slots.append(m | ((i + 1) << 13))
        else:
            slots.append(m | ((y * 8 + x) << 7))
    import oracle as O
    return O.words_of(w[0], slots)


def subtask_masks(subtask):
    """reference Action -> (kind, a, b, goal) via the reference's own nav_utils.get_subtask_obj."""
    ref = H.load_reference()
    ru, nu = ref["recipe_utils"], ref["nav_utils"]
    start, goal = nu.get_subtask_obj(subtask)
    if isinstance(subtask, ru.Chop):
        return (ST_CHOP, H.name_to_mask(start.full_name), 0, H.name_to_mask(goal.full_name))
    if isinstance(subtask, ru.Merge):
        return (ST_MERGE, H.name_to_mask(start[0].full_name), H.name_to_mask(start[1].full_name),
                H.name_to_mask(goal.full_name))
    if isinstance(subtask, ru.Deliver):
        return (ST_DELIVER, H.name_to_mask(start.full_name), 0, H.name_to_mask(goal.full_name))
    raise ValueError(subtask)


def sample_env(level, n_agents, seed, n_steps, eps=0.25):
    """A reference env advanced n_steps with the goal-directed walker (None if it crashed/ended)."""
    rng = np.random.RandomState(seed)
    env = H.make_env(level, n_agents, 100)
    names = env.get_agent_names()
    targets = [None] * n_agents
    for _ in range(n_steps):
        acts = walker_actions(env, rng, eps, targets)
        try:
            with H.quiet():
                _, _, done, _ = env.step({names[i]: DELTA[a] for i, a in enumerate(acts)})
        except (AssertionError, AttributeError):
            return None
        if done:
            return None
    return env


def agent_sets(n_agents):
    names = ["agent-%d" % (i + 1) for i in range(n_agents)]
    out = [((i,), (names[i],)) for i in range(n_agents)]
    out += [((i, j), (names[i], names[j])) for i, j in itertools.combinations(range(n_agents), 2)]
    return out


# ---------------------------------------------------------------------------------------
# lb
# ---------------------------------------------------------------------------------------
def _lb_job(args):
    level, n_agents, seed, n_steps = args[:4]
    levels = args[4] if len(args) > 4 else LEVELS
    from gen_golden import install_custom_level
    install_custom_level(level)
    env = sample_env(level, n_agents, seed, n_steps)
    if env is None:
        return []
    ref = H.load_reference()
    nu = ref["nav_utils"]
    rows = []
    words = pack_env(env)
    for st in env.all_subtasks:
        start, goal = nu.get_subtask_obj(st)
        action_obj = nu.get_subtask_action_obj(st)
        masks = subtask_masks(st)
        for idx, names in agent_sets(n_agents):
            lb = env.get_lower_bound_for_subtask_given_objs(
                subtask=st, subtask_agent_names=names, start_obj=start, goal_obj=goal, subtask_action_obj=action_obj)
            rows.append((list(levels).index(level), n_agents, words, masks, idx[0], idx[1] if len(idx) > 1 else 255, float(lb)))
    return rows


def gen_lb(levels=LEVELS, out_name="lower_bounds.npz", seed0=5000):
    jobs, seed = [], seed0
    for level in levels:
        for n_agents in (1, 2, 3, 4):
            for n_steps in (0, 3, 8, 15, 25, 40, 60):
                seed += 1
                jobs.append((level, n_agents, seed, n_steps, tuple(levels)))
    with Pool(8) as pool:
        res = pool.map(_lb_job, jobs, chunksize=2)
    rows = [r for rr in res for r in rr]
    np.savez_compressed(
        os.path.join(GOLDEN, out_name), levels=np.array(levels),
        level=np.array([r[0] for r in rows], dtype=np.uint8), n_agents=np.array([r[1] for r in rows], dtype=np.uint8),
        state=np.array([r[2] for r in rows], dtype=np.uint32), subtask=np.array([r[3] for r in rows], dtype=np.uint8),
        agent_i=np.array([r[4] for r in rows], dtype=np.uint8), agent_j=np.array([r[5] for r in rows], dtype=np.uint8),
        lb=np.array([r[6] for r in rows], dtype=np.float64))
    lbs = np.array([r[6] for r in rows])
    print("lower bounds:", len(rows), "rows; not-doable (>= perimeter):", int((lbs >= 28).sum()),
          "distinct values:", len(set(lbs.tolist())))


# ---------------------------------------------------------------------------------------
# brtdp
# ---------------------------------------------------------------------------------------
def _brtdp_job(args):
    level, n_agents, seed, n_steps, budget_s = args[:5]
    level1 = len(args) > 5 and args[5]
    env = sample_env(level, n_agents, seed, n_steps)
    if env is None:
        return []
    ref = H.load_reference()
    nu, brtdp = ref["nav_utils"], ref["brtdp"]
    np.random.seed(seed)
    random.seed(seed)
    rows = []
    words = pack_env(env)
    t_start = time.time()
    for st in env.all_subtasks:
        start, goal = nu.get_subtask_obj(st)
        action_obj = nu.get_subtask_action_obj(st)
        masks = subtask_masks(st)
        for idx, names in agent_sets(n_agents):
            if time.time() - t_start > budget_s:
                return rows
            lb = env.get_lower_bound_for_subtask_given_objs(
                subtask=st, subtask_agent_names=names, start_obj=start, goal_obj=goal, subtask_action_obj=action_obj)
            if lb >= env.world.perimeter:  # pruned as not doable in real runs (bd:98-156)
                continue
            planner = brtdp.E2E_BRTDP(alpha=0.01, tau=2, cap=75, main_cap=100)
            others = {}
            if level1:
                # level-1 planning world (e2e_brtdp.py:379-381): any non-empty dict of planners for the
                # other agents; theirs plan the None subtask, i.e. they stay put
                for nm in env.get_agent_names():
                    if nm not in names:
                        op = brtdp.E2E_BRTDP(alpha=0.01, tau=2, cap=75, main_cap=100)
                        with H.quiet():
                            op.set_settings(env=copy.copy(env), subtask=None, subtask_agent_names=(nm,))
                        others[nm] = op
                if not others:
                    continue
            t0 = time.time()
            try:
                with H.quiet():
                    action = planner.get_next_action(env=copy.copy(env), subtask=st, subtask_agent_names=names,
                                                     other_agent_planners=others)
            except (AssertionError, AttributeError, KeyError):
                continue
            key = (planner.cur_state.get_repr(), st)
            v_l, v_u = planner.v_l[key], planner.v_u[key]
            q = np.full(25, np.inf)
            if action is not None:
                with H.quiet():
                    acts = planner.get_actions(state_repr=planner.cur_state.get_repr())
                    for a in acts:
                        qa = planner.Q(state=planner.cur_state, action=a, value_f=planner.v_l)
                        if len(names) == 1:
                            q[DELTA.index(tuple(a))] = qa
                        else:
                            q[5 * DELTA.index(tuple(a[0])) + DELTA.index(tuple(a[1]))] = qa
            rows.append((LEVELS.index(level), n_agents, words, masks, idx[0], idx[1] if len(idx) > 1 else 255,
                         float(lb), float(v_l), float(v_u), q, 1 if action is None else 0, time.time() - t0,
                         len(planner.v_l)))
    return rows


def gen_brtdp(level1=False):
    jobs, seed = [], 9000 + (500 if level1 else 0)
    for level in LEVELS:
        for n_agents in ((2, 3) if not level1 else (2, 3)):
            for n_steps in ((0, 6, 14, 24, 36) if not level1 else (0, 10, 22)):
                seed += 1
                jobs.append((level, n_agents, seed, n_steps, 240.0 if not level1 else 120.0, level1))
    with Pool(8) as pool:
        res = pool.map(_brtdp_job, jobs, chunksize=1)
    rows = [r for rr in res for r in rr]
    np.savez_compressed(
        os.path.join(GOLDEN, "brtdp_values_level1.npz" if level1 else "brtdp_values.npz"), levels=np.array(LEVELS),
        level=np.array([r[0] for r in rows], dtype=np.uint8), n_agents=np.array([r[1] for r in rows], dtype=np.uint8),
        state=np.array([r[2] for r in rows], dtype=np.uint32), subtask=np.array([r[3] for r in rows], dtype=np.uint8),
        agent_i=np.array([r[4] for r in rows], dtype=np.uint8), agent_j=np.array([r[5] for r in rows], dtype=np.uint8),
        lb=np.array([r[6] for r in rows]), v_l=np.array([r[7] for r in rows]), v_u=np.array([r[8] for r in rows]),
        q_l=np.array([r[9] for r in rows]), at_goal=np.array([r[10] for r in rows], dtype=np.uint8),
        seconds=np.array([r[11] for r in rows]), n_states=np.array([r[12] for r in rows], dtype=np.int32))
    conv = sum(1 for r in rows if r[8] - r[7] <= 0.01)
    print("brtdp rows:", len(rows), "converged:", conv, "joint:", sum(1 for r in rows if r[5] != 255),
          "total s:", sum(r[11] for r in rows))


# ---------------------------------------------------------------------------------------
# bd
# ---------------------------------------------------------------------------------------
def _bd_job(args):
    level, n_agents, seed, n_steps, model, observer = args
    env = sample_env(level, n_agents, seed, n_steps)
    if env is None:
        return None
    ref = H.load_reference()
    bd, brtdp = ref["bd"], ref["brtdp"]
    np.random.seed(seed)
    random.seed(seed)
    rng = np.random.RandomState(seed)
    names = env.get_agent_names()
    planner = brtdp.E2E_BRTDP(alpha=0.01, tau=2, cap=75, main_cap=100)
    dele = bd.BayesianDelegator(agent_name=names[observer], all_agent_names=names, model_type=model,
                                planner=planner, none_action_prob=0.5)
    # incomplete subtasks: drop those whose goal object already exists (a crude stand-in for
    # RealAgent.refresh_subtasks; any subset is a legal input for this path)
    incomplete = list(env.all_subtasks)
    with H.quiet():
        dele.set_priors(obs=copy.copy(env), incomplete_subtasks=incomplete, priors_type="uniform")
    if not dele.probs.probs:
        return None
    # arbitrary (seeded) prior so that the fixture is not all-uniform
    keys = dele.probs.enumerate_subtask_allocs()
    pri = rng.rand(len(keys)) + 0.05
    pri /= pri.sum()
    for k, p in zip(keys, pri):
        dele.probs.probs[k] = float(p)
    # one real env step -> obs_tm1 / actions_tm1 exactly as RealAgent.update_subtasks passes them
    targets = [None] * n_agents
    acts = walker_actions(env, rng, 0.3, targets)
    try:
        with H.quiet():
            env.step({names[i]: DELTA[a] for i, a in enumerate(acts)})
    except (AssertionError, AttributeError):
        return None
    calls, soft, rowrec = [], [], []
    real_softmax = bd.sp.special.softmax
    real_pna = dele.prob_nav_actions
    real_q = brtdp.E2E_BRTDP.Q  # patched on the class: get_other_agent_planners copy.copy()s the planner
    q_actions = []
    state_tm1 = pack_env(env.obs_tm1)
    executed = [DELTA.index(tuple(env.agent_actions[nm])) for nm in names]

    def rec_q(self, state, action, value_f):
        if self is planner:
            q_actions.append(action)
        return real_q(self, state=state, action=action, value_f=value_f)

    def act_code(a, joint):
        return 5 * DELTA.index(tuple(a[0])) + DELTA.index(tuple(a[1])) if joint else DELTA.index(tuple(a))

    def rec_softmax(x, *a, **k):
        out = real_softmax(x, *a, **k)
        soft.append((np.array(x, dtype=np.float64), np.array(out, dtype=np.float64)))
        return out

    def rec_pna(**kw):
        n0 = len(soft)
        del q_actions[:]
        p = real_pna(**kw)
        assert len(soft) == n0 + 1
        calls.append((kw["subtask"], tuple(kw["subtask_agent_names"]), float(p), soft[-1]))
        # the row construction, observed: prob_nav_actions calls planner.Q once for the taken action
        # (bd:665) and then once per (filtered) valid action, in order (bd:681-683)
        st, ag = kw["subtask"], tuple(kw["subtask_agent_names"])
        idx = [names.index(a) for a in ag]
        if st is None:
            rowrec.append(dict(kind=0, masks=(0, 0, 0, 0), i=idx[0], j=255, level1=0, valid=[], n_valid=len(soft[-1][0]),
                               act_idx=0 if executed[idx[0]] == 4 else 1))
        else:
            joint = len(ag) == 2
            valid = [act_code(a, joint) for a in q_actions[1:]]
            assert len(valid) == len(soft[-1][0]) and act_code(q_actions[0], joint) in valid
            rowrec.append(dict(kind=2 if joint else 1, masks=subtask_masks(st), i=idx[0], j=idx[1] if joint else 255,
                               level1=int(planner.planner_level == brtdp.PlannerLevel.LEVEL1),
                               valid=valid, n_valid=len(valid), act_idx=valid.index(act_code(q_actions[0], joint))))
        return p

    bd.sp.special.softmax = rec_softmax
    dele.prob_nav_actions = rec_pna
    brtdp.E2E_BRTDP.Q = rec_q
    prior_keys = list(keys)
    prior = dict(dele.probs.probs)
    t0 = time.time()
    try:
        with H.quiet():
            dele.bayes_update(obs_tm1=copy.copy(env.obs_tm1), actions_tm1=env.agent_actions, beta=1.3)
    except (AssertionError, AttributeError, KeyError) as exc:
        return None
    finally:
        bd.sp.special.softmax = real_softmax
        brtdp.E2E_BRTDP.Q = real_q
    post = dict(dele.probs.probs)
    # distinct likelihood rows
    pair_index, pair_rows, rows_out = {}, [], []
    for (st, agents, p, (x, out)), rr in zip(calls, rowrec):
        key = (str(st), agents)
        if key in pair_index:
            continue
        assert abs(out[rr["act_idx"]] - p) < 1e-15  # the softmax entry prob_nav_actions returned (bd:689)
        rows_out.append(rr)
        act_idx = int(np.argmin(np.abs(out - p)))
        assert abs(out[act_idx] - p) < 1e-15
        pair_index[key] = len(pair_rows)
        w = 1 if model == "greedy" else len(agents)
        pair_rows.append((x / 1.3, act_idx, w))
    hyps = []
    for k in prior_keys:
        alive = k in post
        entries = []
        if alive:
            for t in k:
                if model == "greedy" and names[observer] not in t.subtask_agent_names:
                    continue
                entries.append(pair_index[(str(t.subtask), tuple(t.subtask_agent_names))])
        hyps.append((prior[k], alive, entries, post.get(k, 0.0)))
    return dict(level=level, n_agents=n_agents, model=model, observer=observer, hyps=hyps, pairs=pair_rows,
                seconds=time.time() - t0, n_calls=len(calls), rows=rows_out, state=state_tm1, executed=executed,
                subtasks=[subtask_masks(t) for t in env.all_subtasks])


def gen_bd():
    jobs, seed = [], 13000
    for level in ("open-divider_tomato", "partial-divider_tl", "open-divider_salad", "full-divider_salad",
                  "partial-divider_tomato", "full-divider_tl"):
        for (n_agents, model, reps) in ((2, "bd", 4), (2, "dc", 2), (2, "greedy", 2), (2, "up", 1), (3, "bd", 2),
                                        (3, "dc", 1), (4, "dc", 1)):
            for rep in range(reps):
                seed += 1
                jobs.append((level, n_agents, seed, [2, 9, 17, 30][rep % 4], model, seed % n_agents))
    with Pool(8) as pool:
        res = [r for r in pool.map(_bd_job, jobs, chunksize=1) if r is not None]
    n = len(res)
    H_max = max(len(r["hyps"]) for r in res)
    P_max = max(len(r["pairs"]) for r in res)
    A_max = max(max(len(p[0]) for p in r["pairs"]) for r in res)
    E = 4
    prior = np.zeros((n, H_max)); post = np.zeros((n, H_max))
    alive = np.zeros((n, H_max), dtype=np.uint8)
    hyp_pair = np.full((n, H_max, E), 255, dtype=np.uint8)
    pair_w = np.zeros((n, P_max), dtype=np.uint8)
    qdiff = np.zeros((n, P_max, A_max))
    n_valid = np.zeros((n, P_max), dtype=np.uint8)
    act_idx = np.zeros((n, P_max), dtype=np.uint8)
    meta = []
    for r, rec in enumerate(res):
        for h, (p0, al, entries, p1) in enumerate(rec["hyps"]):
            prior[r, h], alive[r, h], post[r, h] = p0, al, p1
            hyp_pair[r, h, :len(entries)] = entries
        for p, (x, ai, w) in enumerate(rec["pairs"]):
            qdiff[r, p, :len(x)] = x
            n_valid[r, p], act_idx[r, p], pair_w[r, p] = len(x), ai, w
        meta.append("%s|%d|%s|%d|H=%d|P=%d" % (rec["level"], rec["n_agents"], rec["model"], rec["observer"],
                                                len(rec["hyps"]), len(rec["pairs"])))
    np.savez_compressed(os.path.join(GOLDEN, "bd_posteriors.npz"), meta=np.array(meta), beta=1.3, prior=prior,
                        posterior=post, alive=alive, hyp_pair=hyp_pair, pair_w=pair_w, qdiff=qdiff,
                        n_valid=n_valid, act_idx=act_idx)
    print("bd posteriors:", n, "updates; H<=%d P<=%d A<=%d; total s %.0f" % (H_max, P_max, A_max,
                                                                                 sum(r["seconds"] for r in res)))
    # the row construction of prob_nav_actions (bd:618-689) for the same calls: obs_tm1, the executed joint
    # action, and per likelihood row the reference's (filtered) valid-action list and taken-action index
    levels = sorted(set(r["level"] for r in res))
    S_max = max(len(r["subtasks"]) for r in res)
    job_level = np.array([levels.index(r["level"]) for r in res], dtype=np.uint8)
    job_agents = np.array([r["n_agents"] for r in res], dtype=np.uint8)
    job_observer = np.array([r["observer"] for r in res], dtype=np.uint8)
    job_model = np.array([r["model"] for r in res])
    job_state = np.array([r["state"] for r in res], dtype=np.uint32)
    job_exec = np.full((n, 4), 4, dtype=np.uint8)
    job_subtasks = np.zeros((n, S_max, 4), dtype=np.uint8)
    job_n_subtasks = np.zeros(n, dtype=np.uint8)
    row_kind = np.zeros((n, P_max), dtype=np.uint8)
    row_masks = np.zeros((n, P_max, 4), dtype=np.uint8)
    row_i = np.full((n, P_max), 255, dtype=np.uint8)
    row_j = np.full((n, P_max), 255, dtype=np.uint8)
    row_level1 = np.zeros((n, P_max), dtype=np.uint8)
    row_valid = np.full((n, P_max, 25), 255, dtype=np.uint8)
    row_n_valid = np.zeros((n, P_max), dtype=np.uint8)
    row_act_idx = np.zeros((n, P_max), dtype=np.uint8)
    for r, rec in enumerate(res):
        job_exec[r, :rec["n_agents"]] = rec["executed"]
        job_n_subtasks[r] = len(rec["subtasks"])
        job_subtasks[r, :len(rec["subtasks"])] = rec["subtasks"]
        assert len(rec["rows"]) == len(rec["pairs"])
        for p, rr in enumerate(rec["rows"]):
            row_kind[r, p], row_masks[r, p], row_i[r, p], row_j[r, p] = rr["kind"], rr["masks"], rr["i"], rr["j"]
            row_level1[r, p], row_n_valid[r, p], row_act_idx[r, p] = rr["level1"], rr["n_valid"], rr["act_idx"]
            row_valid[r, p, :len(rr["valid"])] = rr["valid"]
            assert rr["n_valid"] == n_valid[r, p]
            act_idx[r, p] = rr["act_idx"]  # the true index (the argmin-by-value above picks the first of equal entries)
    np.savez_compressed(os.path.join(GOLDEN, "bd_rows.npz"), levels=np.array(levels), level=job_level,
                        n_agents=job_agents, observer=job_observer, model=job_model, state=job_state,
                        state_layout=np.array("abi2-byte-planes"), executed=job_exec, subtasks=job_subtasks,
                        n_subtasks=job_n_subtasks, row_kind=row_kind, row_masks=row_masks, row_i=row_i, row_j=row_j,
                        row_level1=row_level1, row_valid=row_valid, row_n_valid=row_n_valid, row_act_idx=row_act_idx)
    np.savez_compressed(os.path.join(GOLDEN, "bd_posteriors.npz"), meta=np.array(meta), beta=1.3, prior=prior,
                        posterior=post, alive=alive, hyp_pair=hyp_pair, pair_w=pair_w, qdiff=qdiff,
                        n_valid=n_valid, act_idx=act_idx)
    print("bd rows:", int((row_n_valid > 0).sum()), "rows")


def main(what):
    os.makedirs(GOLDEN, exist_ok=True)
    {"lb": gen_lb, "lb_custom": lambda: gen_lb(("onion-8x8",), "lower_bounds_custom.npz", 7000), "brtdp": gen_brtdp, "brtdp1": lambda: gen_brtdp(level1=True), "bd": gen_bd}[what]()
