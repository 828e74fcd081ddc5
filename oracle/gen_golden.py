"""Generate the committed golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED
REFERENCE (via oracle/ref_harness.py).  Runs only in the build container (needs
/root/reference); the fixtures it writes are what travels to the GPU box.

    python oracle/gen_golden.py env        # tests/golden/env_traces.npz       (path A)
    python oracle/gen_golden.py lb         # tests/golden/lower_bounds.npz     (path B heuristic)
    python oracle/gen_golden.py brtdp      # tests/golden/brtdp_values.npz     (path B values)
    python oracle/gen_golden.py bd         # tests/golden/bd_posteriors.npz    (path C)

Action sequences for path A come from a seeded mixture of uniform-random actions and a
crude goal-directed walker (so that pick / put / chop / merge / deliver all occur); the
walker only chooses ACTIONS - every stored state is the reference's.
"""
import os
import sys
from collections import deque
from multiprocessing import Pool

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, _HERE)
GOLDEN = os.path.join(_ROOT, "tests", "golden")

import ref_harness as H  # noqa: E402

LEVELS = tuple("%s-divider_%s" % (d, r) for d in ("open", "partial", "full") for r in ("tomato", "tl", "salad"))
DELTA = H.ACTIONS
TMAX = 100


# ---------------------------------------------------------------------------------------
# action-sequence generator: drives a live reference env, so it sees true positions
# ---------------------------------------------------------------------------------------
def _floor_bfs(floor, start, blocked):
    dist = {start: (0, None)}
    dq = deque([start])
    while dq:
        c = dq.popleft()
        for a in range(4):
            n = (c[0] + DELTA[a][0], c[1] + DELTA[a][1])
            if n in floor and n not in dist and n not in blocked:
                dist[n] = (dist[c][0] + 1, (c, a))
                dq.append(n)
    return dist


def _first_action(dist, target):
    a = None
    c = target
    while dist[c][1] is not None:
        c, a = dist[c][1]
    return a


def walker_actions(env, rng, eps, targets):
    """One joint action; `targets` is per-agent mutable state (square to interact with)."""
    ref = H.load_reference()
    core = ref["core"]
    floor = {o.location for o in env.world.objects.get("Floor", [])}
    squares = [o for o in env.world.get_object_list() if isinstance(o, core.GridSquare) and o.collidable]
    agent_locs = {a.location for a in env.sim_agents}
    acts = []
    for i, ag in enumerate(env.sim_agents):
        if rng.rand() < eps:
            acts.append(int(rng.randint(0, 5)))
            continue
        tgt = targets[i]
        if tgt is None or rng.rand() < 0.05:
            # choose a square: prefer something useful for what the agent holds
            objs = [o for o in env.world.get_object_list() if isinstance(o, core.Object) and not o.is_held]
            cand = []
            if ag.holding is None:
                cand = [o.location for o in objs]
            else:
                h = ag.holding
                if h.needs_chopped():
                    cand = [s.location for s in squares if isinstance(s, core.Cutboard)]
                elif h.is_deliverable() and rng.rand() < 0.7:
                    cand = [s.location for s in squares if isinstance(s, core.Delivery)]
                else:
                    cand = [o.location for o in objs if core.mergeable(h, o)]
            if not cand or rng.rand() < 0.25:
                cand = [s.location for s in squares]
            tgt = cand[rng.randint(len(cand))]
            targets[i] = tgt
        # floor cells adjacent to the target square
        dist = _floor_bfs(floor, ag.location, agent_locs - {ag.location})
        best = None
        for a in range(4):
            f = (tgt[0] - DELTA[a][0], tgt[1] - DELTA[a][1])
            if f in dist and (best is None or dist[f][0] < best[0]):
                best = (dist[f][0], f, a)
        if best is None:
            targets[i] = None
            acts.append(int(rng.randint(0, 5)))
        elif best[0] == 0:
            acts.append(best[2])  # face the square: interact
            targets[i] = None
        else:
            acts.append(_first_action(dist, best[1]))
    return acts


CUSTOM_LEVEL_DIR = os.path.join(GOLDEN, "levels")


def install_custom_level(level):
    """Levels that are not among the reference's nine live under tests/golden/levels/; the reference
    opens `utils/levels/<name>.txt` relative to its cwd (env:146), i.e. inside the scratch copy."""
    src = os.path.join(CUSTOM_LEVEL_DIR, level + ".txt")
    if os.path.exists(src):
        import shutil
        shutil.copy(src, os.path.join(H.load_reference()["scratch"], "utils", "levels", level + ".txt"))


def gen_one_trace(args):
    level, n_agents, seed, eps, max_t = args
    install_custom_level(level)
    rng = np.random.RandomState(seed)
    env = H.make_env(level, n_agents, max_t)
    names = env.get_agent_names()
    targets = [None] * n_agents
    rec = dict(level=level, n_agents=n_agents, max_t=max_t, seed=seed, eps=eps,
               actions=[], states=[H.canonical(env)], done=[False], reward=[0], ncoll=[0],
               executed=[[4] * n_agents], crashed=False)
    for _ in range(TMAX):
        acts = walker_actions(env, rng, eps, targets)
        ad = {names[i]: DELTA[a] for i, a in enumerate(acts)}
        nc0 = len(env.collisions)
        try:
            with H.quiet():
                _, reward, done, _info = env.step(ad)
        except (AssertionError, AttributeError):  # world.py:417 (its message itself raises)
            rec["crashed"] = True  # co-located agents both holding
            break
        rec["actions"].append(acts)
        rec["states"].append(H.canonical(env))
        rec["done"].append(bool(done))
        rec["reward"].append(int(reward))
        rec["ncoll"].append(len(env.collisions) - nc0)
        rec["executed"].append([DELTA.index(tuple(env.agent_actions[nm])) for nm in names])
        if done:
            break
    return rec


def gen_env(levels=LEVELS, out_name="env_traces.npz", seed0=1000):
    jobs = []
    seed = seed0
    for level in levels:
        for n_agents in (1, 2, 3, 4):
            # eps = 1.0 is pure uniform-random (cfg-2 style); small eps is goal-directed
            for eps in (1.0, 1.0, 0.5, 0.3, 0.3, 0.15, 0.15, 0.15, 0.05, 0.05):
                seed += 1
                jobs.append((level, n_agents, seed, eps, 100))
            for max_t in (7, 23):
                seed += 1
                jobs.append((level, n_agents, seed, 0.3, max_t))
    with Pool(8) as pool:
        recs = pool.map(gen_one_trace, jobs, chunksize=4)
    n = len(recs)
    actions = np.full((n, TMAX, 4), 4, dtype=np.uint8)
    executed = np.full((n, TMAX + 1, 4), 4, dtype=np.uint8)
    t_arr = np.zeros((n, TMAX + 1), dtype=np.uint8)
    done = np.zeros((n, TMAX + 1), dtype=np.uint8)
    reward = np.zeros((n, TMAX + 1), dtype=np.uint8)
    ncoll = np.zeros((n, TMAX + 1), dtype=np.uint8)
    agents = np.zeros((n, TMAX + 1, 4, 2), dtype=np.uint8)
    keys = np.full((n, TMAX + 1, 6), 0x3FFF, dtype=np.uint16)
    length = np.zeros(n, dtype=np.int32)
    meta = np.zeros((n, 4), dtype=np.int32)  # level idx, n_agents, max_t, crashed
    stats = dict(deliver=0, success=0, crashed=0, merged=0, chopped=0)
    for r, rec in enumerate(recs):
        L = len(rec["actions"])
        length[r] = L
        meta[r] = (list(levels).index(rec["level"]), rec["n_agents"], rec["max_t"], int(rec["crashed"]))
        stats["crashed"] += int(rec["crashed"])
        if L:
            actions[r, :L, :rec["n_agents"]] = np.array(rec["actions"], dtype=np.uint8)
        for s in range(L + 1):
            t, ags, items = rec["states"][s]
            t_arr[r, s] = t
            done[r, s] = rec["done"][s]
            reward[r, s] = rec["reward"][s]
            ncoll[r, s] = rec["ncoll"][s]
            executed[r, s, :rec["n_agents"]] = rec["executed"][s]
            for i, (x, y, hm) in enumerate(ags):
                agents[r, s, i] = (y * 8 + x, hm)
            ks = sorted((m << 7) | ((y * 8 + x) << 1) | held for (m, x, y, held) in items)
            keys[r, s, :len(ks)] = ks
        last_items = rec["states"][L][2]
        stats["success"] += int(rec["reward"][L])
        stats["merged"] += int(any(bin(m & 15).count("1") > 1 for (m, _, _, _) in last_items))
        stats["chopped"] += int(any(m >> 4 for (m, _, _, _) in last_items))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, out_name), levels=np.array(levels), meta=meta,
                        length=length, actions=actions, executed=executed, t=t_arr, done=done,
                        reward=reward, ncoll=ncoll, agents=agents, keys=keys)
    print("env traces:", n, "steps:", int(length.sum()), stats)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "env"
    if what == "env":
        gen_env()
    elif what == "env_custom":  # an 8x8 kitchen with an onion, three plates (6 objects) and OnionSalad
        gen_env(levels=("onion-8x8",), out_name="env_traces_custom.npz", seed0=3000)
    else:
        import gen_golden_plan  # noqa: F401  (path B / C generators live there)
        gen_golden_plan.main(what)
