"""CPU timings of the UNMODIFIED Python reference on its own hot paths,
taken in the build container - the only place /root/reference exists; the GPU box cannot run it.
Writes profiles/r02_python_reference_cpu.json, which bench.py attaches to `cpu_baseline` beside the
C port it times on the GPU box's own cores.

    python oracle/time_reference.py [--box-seconds 600]

C1  main_loop of config 1 (2 agents, open-divider_salad, bd/bd, seed 1), single process, time-boxed:
    env steps completed, bayes_update calls -> agent-steps/s, posterior updates/s; plus the complete
    open-divider_tomato bd/bd episode.
C2  the same loop over seeds 1..8 in a multiprocessing.Pool(8), time-boxed.
C3  env.step only: 100 trajectories of config 2 (2 agents, partial-divider_tl, 100 uniform-random steps),
    single process and Pool(8).
C5  BayesianDelegator.bayes_update alone on sampled states (the calls of gen_golden_plan's bd fixture).
(C4, the BRTDP runs, are the `seconds` column of tests/golden/brtdp_values.npz: 14 030 CPU-seconds for 409
(state, subtask, agents) problems.)
"""
import argparse
import json
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import ref_harness as H  # noqa: E402


def _main_loop(args):
    """The body of main.main_loop (main.py:85-117) with counters and a wall-clock box; the Bag is left out
    (its directory is a hard-coded Windows path, metrics_bag.py:9)."""
    level, models, seed, box = args
    import random
    ref = H.load_reference()
    arglist = H.make_arglist(level, len(models), 100, models=list(models), seed=seed)
    np.random.seed(seed)
    random.seed(seed)
    counts = {"bayes": 0}
    real = ref["bd"].BayesianDelegator.bayes_update

    def counted(self, *a, **k):
        counts["bayes"] += 1
        return real(self, *a, **k)

    ref["bd"].BayesianDelegator.bayes_update = counted
    t0 = time.time()
    steps, done = 0, False
    try:
        with H.quiet():
            env = ref["env"].OvercookedEnvironment(arglist)
            obs = env.reset()
            env.game = H._StubGame()
            agents = [ref["agent"].RealAgent(arglist=arglist, name="agent-%d" % (i + 1), id_color=ref["agent"].COLORS[i],
                                             recipes=env.recipes) for i in range(len(models))]
            while not env.done() and time.time() - t0 < box:
                action_dict = {a.name: a.select_action(obs=obs) for a in agents}
                obs, reward, done, info = env.step(action_dict=action_dict)
                for a in agents:
                    a.refresh_subtasks(world=env.world)
                steps += 1
    finally:
        ref["bd"].BayesianDelegator.bayes_update = real
    dt = time.time() - t0
    return dict(level=level, models=list(models), seed=seed, env_steps=steps, seconds=dt, finished=bool(env.done()),
                successful=bool(getattr(env, "successful", False)), bayes_updates=counts["bayes"],
                agent_steps_per_sec=steps * len(models) / dt, posterior_updates_per_sec=counts["bayes"] / dt)


def _env_steps(args):
    seed, n_traj = args
    rng = np.random.RandomState(seed)
    total, dt = 0, 0.0
    for _ in range(n_traj):
        env = H.make_env("partial-divider_tl", 2, 100)
        names = env.get_agent_names()
        acts = rng.randint(0, 5, size=(100, 2))
        t0 = time.time()
        with H.quiet():
            for a in acts:
                try:
                    _, _, done, _ = env.step({names[i]: H.ACTIONS[a[i]] for i in range(2)})
                except (AssertionError, AttributeError):
                    break
                total += 1
                if done:
                    break
        dt += time.time() - t0
    return total, dt


def _bayes_only(seed):
    """one set_priors + one env step + one bayes_update on a sampled state, timed separately"""
    import copy
    from gen_golden_plan import sample_env, walker_actions, DELTA
    ref = H.load_reference()
    env = sample_env("open-divider_salad", 2, seed, [2, 9, 17][seed % 3])
    if env is None:
        return None
    names = env.get_agent_names()
    planner = ref["brtdp"].E2E_BRTDP(alpha=0.01, tau=2, cap=75, main_cap=100)
    dele = ref["bd"].BayesianDelegator(agent_name=names[0], all_agent_names=names, model_type="bd", planner=planner,
                                       none_action_prob=0.5)
    with H.quiet():
        dele.set_priors(obs=copy.copy(env), incomplete_subtasks=list(env.all_subtasks), priors_type="uniform")
        rng = np.random.RandomState(seed)
        acts = walker_actions(env, rng, 0.3, [None, None])
        try:
            env.step({names[i]: DELTA[a] for i, a in enumerate(acts)})
        except (AssertionError, AttributeError):
            return None
        t0 = time.time()
        try:
            dele.bayes_update(obs_tm1=copy.copy(env.obs_tm1), actions_tm1=env.agent_actions, beta=1.3)
        except (AssertionError, AttributeError, KeyError):
            return None
    return time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--box-seconds", type=float, default=600.0)
    args = ap.parse_args()
    cores = os.cpu_count() or 1
    out = {"where": "build container (the GPU box has no /root/reference)", "cores": cores,
           "python": sys.version.split()[0], "box_seconds": args.box_seconds}
    # C3
    n, dt = _env_steps((1, 100))
    out["C3_env_step_1core"] = {"env_steps": n, "seconds": dt, "agent_steps_per_sec": 2 * n / dt}
    t0 = time.time()
    with Pool(8) as pool:
        res = pool.map(_env_steps, [(s, 100) for s in range(1, 9)])
    wall = time.time() - t0
    n8 = sum(r[0] for r in res)
    out["C3_env_step_pool8"] = {"env_steps": n8, "wall_seconds": wall, "agent_steps_per_sec": 2 * n8 / wall,
                                "note": "wall time of the pool, resets and process start included"}
    print(json.dumps(out, indent=1), flush=True)
    # C5
    with Pool(8) as pool:
        ts = [t for t in pool.map(_bayes_only, range(100, 148)) if t is not None]
    out["C5_bayes_update"] = {"calls": len(ts), "median_seconds": float(np.median(ts)), "mean_seconds": float(np.mean(ts)),
                              "posterior_updates_per_sec_per_core": 1.0 / float(np.mean(ts)),
                              "note": "bayes_update alone (likelihoods from the planner's heuristic v_l: no BRTDP run inside)"}
    print(json.dumps(out["C5_bayes_update"], indent=1), flush=True)
    # C1 (+ the finished-episode datum) and C2 run side by side: 2 + 8 processes on 8 cores would distort C1,
    # so C1 first, then C2
    with Pool(2) as pool:
        c1, c1t = pool.map(_main_loop, [("open-divider_salad", ("bd", "bd"), 1, args.box_seconds),
                                        ("open-divider_tomato", ("bd", "bd"), 1, 3 * args.box_seconds)])
    out["C1_main_loop_cfg1"] = c1
    out["C1_finished_episode_tomato"] = c1t
    print(json.dumps({"C1": c1, "C1t": c1t}, indent=1), flush=True)
    t0 = time.time()
    with Pool(8) as pool:
        res = pool.map(_main_loop, [("open-divider_salad", ("bd", "bd"), s, args.box_seconds) for s in range(1, 9)])
    wall = time.time() - t0
    out["C2_main_loop_pool8"] = {"wall_seconds": wall, "env_steps": sum(r["env_steps"] for r in res),
                                 "bayes_updates": sum(r["bayes_updates"] for r in res),
                                 "agent_steps_per_sec": 2 * sum(r["env_steps"] for r in res) / wall,
                                 "posterior_updates_per_sec": sum(r["bayes_updates"] for r in res) / wall,
                                 "finished": sum(r["finished"] for r in res)}
    path = os.path.join(_HERE, "..", "profiles", "r02_python_reference_cpu.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
