#!/usr/bin/env python
"""bench.py - headline benchmark of gym-cooking_b200 (contract: see the task prompt / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W            # our arm (N > 1: under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

Workload = BASELINE.json configs[1] ("cfg-2"): 2 agents, partial-divider_tl, 2^20 envs per GPU,
uniform-random actions (philox stream).  A *step* is one env.step over one
batch of 2^20 envs = one gc_env_step launch.  To keep the timed region out of the 126 MB L2 the
launches cycle through a ring of 16 independent batches per GPU (16 x (16 MB state + 2 MB
actions + 1 MB reward/done) = 304 MB), so every launch streams its inputs from HBM.
`value` = agent-steps/s with everything resident in HBM; `e2e` = the same metric through the
public API (OvercookedEnvironment.step) with the step's actions copied from pinned host memory
and its reward/done bytes copied back, every step.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LEVEL, N_AGENTS, N_ENVS, HORIZON = "partial-divider_tl", 2, 1 << 20, 100
RING = 16
SEED = 1234
METRIC, UNIT = "agent_steps_per_sec", "agent-steps/s"
# algorithmic bytes per env-step of gc_env_step: 16 B state read + 16 B
# state write + n_agents action bytes + 1 reward/done byte
BYTES_PER_ENV_STEP = 33 + N_AGENTS


def workload_config(n_gpus):
    return {
        "workload": "cfg-2: 2-agent partial-divider_tl, 2^20 envs per GPU, uniform-random actions, env step only",
        "level": LEVEL, "n_agents": N_AGENTS, "envs_per_gpu": N_ENVS, "horizon": HORIZON,
        "actions": "philox4x32-10(seed=1234, ctr=(t, env)) -> mulhi(word, 5), resident in HBM",
        "l2": "ring of %d batches per GPU (%d MB > 126 MB L2): inputs larger than L2" % (
            RING, RING * N_ENVS * (16 + N_AGENTS + 1) >> 20),
        "parallelism": "env-batch sharding x%d, no data-path collective" % n_gpus,
    }


# ----------------------------------------------------------------------------------------
# clocks: sample NVML during the timed region
# ----------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit and name != "gpu_idle":
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            time.sleep(0.005)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self.sample()
        self._stop.set()
        if self._thread:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------
# CPU legs (oracle port): cpu_baseline of our arm and the whole reference arm
# ----------------------------------------------------------------------------------------
def cpu_port_rate(n_envs, n_steps, n_threads):
    """agent-steps/s of the CPU oracle (oracle/gc_oracle.c) on cfg-2's own action stream."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import gym_cooking_b200 as gcb
    lv = O.parse_level(gcb.levels.level_text(LEVEL), HORIZON)
    st = O.reset_state(lv, N_AGENTS, n_envs)
    t0 = time.perf_counter()
    O.rollout_batch(lv, st, N_AGENTS, n_steps, seed=SEED, n_threads=n_threads)
    dt = time.perf_counter() - t0
    return n_envs * n_steps * N_AGENTS / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # only rank 0 runs the CPU arm
    cores = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import gym_cooking_b200 as gcb
    lv = O.parse_level(gcb.levels.level_text(LEVEL), HORIZON)
    st = O.reset_state(lv, N_AGENTS, N_ENVS)
    acts = O.fill_actions(N_ENVS, N_AGENTS, 8, seed=SEED)  # the same philox stream our arm steps with
    k = 0
    for _ in range(args.warmup):
        O.step_batch(lv, st, acts[k % 8], N_AGENTS, n_threads=cores, want_collisions=False)
        k += 1
    t0 = time.perf_counter()
    for s in range(args.steps):
        if k % HORIZON == 0:
            st = O.reset_state(lv, N_AGENTS, N_ENVS)
        O.step_batch(lv, st, acts[k % 8], N_AGENTS, n_threads=cores, want_collisions=False)
        k += 1
    dt = time.perf_counter() - t0
    value = args.steps * N_ENVS * N_AGENTS / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "every step = one env.step over 2^20 envs on %d host threads "
                                   "(C port of the reference's pure-Python step; the Python reference "
                                   "itself measured 1.3e3 agent-steps/s per core, BASELINE.md)" % cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import gym_cooking_b200 as gcb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident workload: RING independent 2^20-env batches with their action streams ----
    ring = [gcb.KitchenBatch(LEVEL, N_AGENTS, N_ENVS, HORIZON, device=dev) for _ in range(RING)]
    env0 = rank * RING * N_ENVS  # global env indices: results do not depend on the GPU count
    actions = [kb.random_actions(HORIZON, env0=env0 + r * N_ENVS, seed=SEED) for r, kb in enumerate(ring)]
    # per-step action views made once: at ~8 us per kernel the Python side of a step must stay well below
    # that, or a launch loop measures the interpreter instead of the GPU
    views = [[actions[r][t] for t in range(HORIZON)] for r in range(RING)]
    PERIOD = RING * HORIZON

    # Step k of the run (k = 0, 1, ...) advances batch k % RING with the actions of episode time
    # (k // RING) % HORIZON; a batch starts new episodes (gc_env_reset) when its episode time wraps.
    def do_step(k):
        r, t = k % RING, (k // RING) % HORIZON
        n = 1
        if t == 0 and k >= RING:
            ring[r].reset()  # gc_env_reset (the reward/done memset is torch's, not counted)
            n = 2
        ring[r].step(views[r][t])
        return n

    k = 0
    for _ in range(max(args.warmup, 3)):
        do_step(k)
        k += 1
    # A plain Python loop issuing one launch per step is at the edge of keeping the GPU fed (the kernel takes
    # ~8 us), so the timed steps are replayed as CUDA graphs - the same launches in the same order - whatever
    # --steps is: chunks of GRAPH_STEPS consecutive steps aligned to the run's period (4 reusable graphs) plus
    # one graph each for a ragged head / tail.  Graphs are captured and uploaded (cudaGraphUpload) before the
    # timed region.  GC_BENCH_NO_GRAPH=1 times the plain loop instead; the other mode's rate is measured
    # right after the timed region and reported as config.launch_ab.
    GRAPH_STEPS = 400
    use_graphs = os.environ.get("GC_BENCH_NO_GRAPH") is None
    side = torch.cuda.Stream(device=dev)
    graph_cache = {}

    def capture(k_lo, k_hi):
        key = (k_lo % PERIOD, k_hi - k_lo, k_lo >= RING)
        if key not in graph_cache:
            g = torch.cuda.CUDAGraph()
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    n = sum(do_step(kk) for kk in range(k_lo, k_hi))
            torch.cuda.current_stream(dev).wait_stream(side)
            graph_upload(g, torch.cuda.current_stream(dev))
            graph_cache[key] = (g, n)
        return graph_cache[key]

    def plan_for(k_lo, k_hi):
        plan, kk = [], k_lo
        while kk < k_hi:
            nxt = min(k_hi, (kk // GRAPH_STEPS + 1) * GRAPH_STEPS)
            plan.append(capture(kk, nxt))
            kk = nxt
        return plan

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1), n

    def run_graphs(plan):
        n = 0
        for g, n_launch in plan:
            g.replay()
            n += n_launch
        return n

    def run_loop(k_lo, k_hi):
        return sum(do_step(kk) for kk in range(k_lo, k_hi))

    # capturing a graph does not execute it: the states are where the warm-up left them
    plan = plan_for(k, k + args.steps) if use_graphs else None
    sampler = ClockSampler(local)
    sampler.start()
    if use_graphs:
        ms, timed_launches = timed(lambda: run_graphs(plan))
    else:
        ms, timed_launches = timed(lambda: run_loop(k, k + args.steps))
    clocks = sampler.stop()
    k += args.steps
    # the other launch mode, for the A/B line: at least 400 steps (the first dozens of launches of a host loop
    # are not pipelined yet; 20 steps would measure that ramp), at most 2000
    ab_steps = min(max(args.steps, 400), 2000)
    if use_graphs:
        ab_ms, _ = timed(lambda: run_loop(k, k + ab_steps))
    else:
        ab_plan = plan_for(k, k + ab_steps)
        ab_ms, _ = timed(lambda: run_graphs(ab_plan))
    k += ab_steps
    if world > 1:
        t = torch.tensor([ms, ab_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ab_ms = (float(v) for v in t.tolist())
    value = world * args.steps * N_ENVS * N_AGENTS / (ms * 1e-3)

    # ---- end to end through the public API: host actions in, reward/done bits out, per step ----
    # Two host-side action formats through the same call, OvercookedEnvironment.step(pinned host tensor):
    # one joint index per env (uint8[N] = 5 * a_1 + a_2, the planners' joint-action index: 1 MB per step)
    # - the headline e2e - and the per-agent bytes (uint8[N][2]: 2 MB per step).
    ns = argparse.Namespace(level=LEVEL, num_agents=N_AGENTS, max_num_timesteps=HORIZON, max_num_subtasks=14,
                            seed=1, model1=None, model2=None, model3=None, model4=None)
    env = gcb.OvercookedEnvironment(ns, num_envs=N_ENVS, device=dev, track_collisions=False)
    byte_actions = [actions[0][s].cpu().pin_memory() for s in range(8)]
    joint_actions = [(actions[0][s][:, 0] * 5 + actions[0][s][:, 1]).to(torch.uint8).cpu().pin_memory() for s in range(8)]
    # at least 400 timed steps whatever --steps is (28 ms): a 20-step e2e leg would sit inside that ramp
    e2e_steps = min(max(args.steps, 400), 2000)
    # warm-up: one untimed block of the same length.  The first few hundred host-driven steps run up to
    # 30-45 % slower than the steady state (PCIe link / copy-engine / host ramp; scripts/e2e_numa_probe.py:
    # blocks of 400 steps measure 1.3-1.8e10, then 2.4e10 for every later block), so 3 steps are not enough
    e2e_warmup = max(args.warmup, 3, e2e_steps)

    def e2e_run(host_actions):
        env.reset()
        for s in range(e2e_warmup):
            if s and s % HORIZON == 0:
                env.reset()
            env.step(host_actions[s % 8])
        env.reset()
        barrier()
        t0 = time.perf_counter()
        for s in range(e2e_steps):
            if s and s % HORIZON == 0:
                env.reset()
            env.step(host_actions[s % 8])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * e2e_steps * N_ENVS * N_AGENTS / dt

    e2e_bytes_value = e2e_run(byte_actions)
    e2e_value = e2e_run(joint_actions)

    # the same steps, vector-env style: two batches of 2^20 envs stepped alternately with step_async / step_wait
    # (each on its own stream), so that one batch's PCIe copies overlap the other's kernel.  Every step still
    # moves its actions in and its results out; reported beside the synchronous number, not instead of it.
    env_b = gcb.OvercookedEnvironment(ns, num_envs=N_ENVS, device=dev, track_collisions=False)

    def e2e_async_run(host_actions):
        pair = (env, env_b)
        for e_ in pair:
            e_.reset()
        torch.cuda.synchronize()
        timing = None
        for phase in ("warm", "timed"):
            if phase == "timed":
                for e_ in pair:
                    e_.reset()
                barrier()
                t0 = time.perf_counter()
            pair[0].step_async(host_actions[0])
            for s in range(1, 2 * e2e_steps):
                if s % (2 * HORIZON) < 2 and s >= 2 * HORIZON:
                    pair[s % 2].reset()
                pair[s % 2].step_async(host_actions[s % 8])
                pair[(s - 1) % 2].step_wait()
            pair[(2 * e2e_steps - 1) % 2].step_wait()
            torch.cuda.synchronize()
            if phase == "timed":
                timing = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([timing], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            timing = float(t.item())
        return world * 2 * e2e_steps * N_ENVS * N_AGENTS / timing

    e2e_async_value = e2e_async_run(joint_actions)
    del env_b

    # ---- the one collective of this path: reduce the episode statistics over ranks ----
    stats = torch.zeros(gcb._lib.STATS_LEN, dtype=torch.int64, device=dev)
    for kb in ring:
        kb.stats(out=stats)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)  # NCCL over NVLink, 1064 bytes
    stats = stats.cpu().tolist()

    # ---- cfg-5: four agents, mixed bd/up/dc/fb/greedy, all nine levels; every rank runs its own shard of episodes
    # (no data-path collective), the totals are reduced like the episode statistics ----
    cfg5 = None
    if not args.no_secondary and args.cfg5_envs > 0:
        from gym_cooking_b200 import batched_agents
        r5 = batched_agents.run_mixed(args.cfg5_envs, 4, shard=rank, device=dev)
        from gym_cooking_b200 import sharding
        t5 = sharding.reduce_mixed_totals(r5, device=dev)  # counts summed over ranks, time of the slowest rank
        tot = [t5[k] for k in sharding.MIXED_TOTALS]
        tmax = t5["seconds"]
        cfg5 = {"metric": "cfg5_agent_steps_per_sec", "value": tot[1] / tmax, "unit": "agent-steps/s",
                "config": "cfg-5: 4 agents, model types bd/up/dc/fb/greedy rotated over seats and levels, all nine levels, "
                          "%d envs per level per GPU x %d GPU(s), horizon 100, full delegation loop from reset with a cold "
                          "planner memo (max over ranks of the summed per-level wall time)" % (args.cfg5_envs, world),
                "posterior_updates_per_sec": tot[2] / tmax, "seconds": tmax, "envs": tot[0], "delivered": tot[3],
                "planning_states_solved": tot[4], "planner_lookups": tot[5], "completed_subtasks": tot[6],
                "note": "planner-bound: with four agents nearly every env-step reaches a planning state nobody has seen, "
                        "and each such state costs 20 x n_subtasks exact searches (DESIGN.md section 7)",
                "per_level_rank0": [{k_: v_ for k_, v_ in rec.items()} for rec in r5["per_level"]]}
    secondary = secondary_metrics(gcb, torch, dev) if (rank == 0 and not args.no_secondary) else None
    if secondary is not None and cfg5 is not None:
        secondary.append(cfg5)
    if rank == 0:
        peak, peak_src = hbm_peak_gbs()
        launch_s = ms * 1e-3 / args.steps  # average gc_env_step launch, launch gaps included
        achieved = BYTES_PER_ENV_STEP * N_ENVS / launch_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
        traffic_note = None
        if os.path.exists(tpath):  # from the committed ncu capture of this kernel, not measured in this run
            tj = json.load(open(tpath))
            traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("note")
        cpu_value, cpu_dt = cpu_port_rate(1 << 18, HORIZON, 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            # `config` is the workload only (identical in both arms); how the steps were launched is its own key
            "config": workload_config(world),
            "launch": {"mode": ("CUDA graphs of up to %d consecutive steps (captured and uploaded before the timed region)"
                                % GRAPH_STEPS if use_graphs else "one gc_env_step call per step from a Python loop"),
                       "ab": {"mode": "python loop, one gc_env_step call per step" if use_graphs else "CUDA graphs",
                              "steps": ab_steps, "us_per_step": ab_ms * 1e3 / ab_steps,
                              "value": world * ab_steps * N_ENVS * N_AGENTS / (ab_ms * 1e-3)}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_ENVS,
                    "d2h_bytes_per_step": (N_ENVS + 31) // 32 * 8,
                    "results": "done / reward bit planes (2 bits per env)",
                    "steps": e2e_steps, "warmup_steps": e2e_warmup,
                    "api": "OvercookedEnvironment(arglist, num_envs=2^20).step(pinned uint8[N]): one joint action "
                           "index per env (5 * a_1 + a_2)",
                    "per_agent_bytes": {"value": e2e_bytes_value, "h2d_bytes_per_step": N_ENVS * N_AGENTS,
                                        "api": "the same call with pinned uint8[N][2] (one byte per agent)"},
                    "async_two_batches": {"value": e2e_async_value, "h2d_bytes_per_step": N_ENVS,
                                          "d2h_bytes_per_step": (N_ENVS + 31) // 32 * 8,
                                          "api": "two batches of 2^20 envs alternating step_async(pinned uint8[N]) / "
                                                 "step_wait() (vector-env style), each on its own stream"}},
            "gpu_launches": timed_launches,
            "roofline": {"bound": "hbm", "kernel": "step2_kernel<NA=2,NOBJ=4,EXTRAS=0,BITS=0,MULTI=0> (gc_env_step, plain step)", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "bytes_per_launch": BYTES_PER_ENV_STEP * N_ENVS, "launch_us": launch_s * 1e6,
                         "traffic": traffic, "traffic_source": traffic_note},
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "2^18 envs x 100 steps of the same action stream, %.1f s" % cpu_dt,
                             "python_reference": python_reference_numbers()},
            "secondary": secondary,
            "episode_stats": {"episodes": stats[0], "successes": stats[1], "sum_t_done": stats[2],
                              "running": stats[4], "reduced_with": "nccl all_reduce" if world > 1 else "single rank"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def graph_upload(g, stream):
    """cudaGraphUpload: move the first-launch cost of a captured graph out of the timed region."""
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaGraphUpload.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        return rt.cudaGraphUpload(g.raw_cuda_graph_exec(), stream.cuda_stream) == 0
    except Exception:
        return False


def python_reference_numbers():
    """Timings of the unmodified Python reference (main.py single process and a multiprocessing pool, env.step
    alone, bayes_update alone) taken by oracle/time_reference.py in the build container - the reference cannot
    travel to the GPU box, so these are attached from profiles/, not measured in this run."""
    path = os.path.join(ROOT, "profiles", "r02_python_reference_cpu.json")
    if not os.path.exists(path):
        return None
    d = json.load(open(path))
    pick = lambda k, *f: {x: d[k][x] for x in f if x in d.get(k, {})}
    return {"kind": "python-reference", "where": d.get("where"), "cores": d.get("cores"),
            "env_step_1core": pick("C3_env_step_1core", "agent_steps_per_sec"),
            "env_step_pool8": pick("C3_env_step_pool8", "agent_steps_per_sec"),
            "main_py_cfg1_single_process": pick("C1_main_loop_cfg1", "agent_steps_per_sec", "posterior_updates_per_sec",
                                                "env_steps", "seconds", "finished"),
            "main_py_tomato_episode": pick("C1_finished_episode_tomato", "agent_steps_per_sec", "env_steps", "seconds",
                                           "finished"),
            "main_py_cfg1_pool8": pick("C2_main_loop_pool8", "agent_steps_per_sec", "posterior_updates_per_sec",
                                       "wall_seconds"),
            "bayes_update_per_core": pick("C5_bayes_update", "posterior_updates_per_sec_per_core")}


def hbm_peak_gbs():
    """(GB/s, source): the driver-measured copy bandwidth, else the profiling recipe's fallback"""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def secondary_metrics(gcb, torch, dev):
    """The other two hot paths of BASELINE.json's metric at the sizes SURVEY 8d states: BD posterior updates/s
    (cfg-4 shape, CUDA events, 3 warm-up + 10 timed launches), the whole delegation loop of cfg-4 (2^18 envs x 100)
    and the planner on cfg-3 (2^20 envs)."""
    import itertools

    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    out = []
    # path C: 2 agents open-divider_salad has H = 8 hypotheses over P = 8 likelihood rows of A <= 5 actions
    n, H, P, A, E = 1 << 21, 8, 8, 5, 2  # 2^21 rows x 248 B = 520 MB > L2
    g = torch.Generator(device=dev).manual_seed(7)
    probs = torch.rand((n, H), device=dev, generator=g)
    hyp = torch.randint(0, P, (n, H, E), device=dev, generator=g, dtype=torch.uint8)
    w = torch.randint(1, 3, (n, P), device=dev, generator=g, dtype=torch.uint8)
    qd = torch.randn((n, P, A), device=dev, generator=g)
    nv = torch.full((n, P), A, device=dev, dtype=torch.uint8)
    ai = torch.randint(0, A, (n, P), device=dev, generator=g, dtype=torch.uint8)
    t = timed(lambda: gcb.bd_posterior(probs, None, hyp, w, qd, nv, ai, 1.3), 10)
    bytes_per = 8 * H + 4 * P * A + 3 * P + H * E
    out.append({"metric": "bd_posterior_updates_per_sec", "value": n / t, "unit": "updates/s", "dtype": "f32",
                "config": "cfg-4 shape: H=8 hypotheses, P=8 likelihood rows, A=5 actions, 2^21 rows resident (520 MB)",
                "bytes_per_update": bytes_per, "achieved_gbs": n * bytes_per / t / 1e9,
                "roofline_frac": n * bytes_per / t / 1e9 / hbm_peak_gbs()[0]})
    del probs, hyp, w, qd, nv, ai
    # cfg-5, env half: 4 agents, per-env level id uniform over the nine levels, 2^20 envs, philox actions
    # (seed 1236) - step2_kernel<4,4,EXTRAS=0,MULTI=1>; ring of 4 batches (80 MB of state + actions per step)
    n5, ring5 = 1 << 20, 4
    g5 = torch.Generator(device="cpu").manual_seed(1236)
    kbs = []
    for r in range(ring5):
        lid = torch.randint(0, 9, (n5,), generator=g5, dtype=torch.uint8)
        kbs.append(gcb.KitchenBatch(list(gcb.levels.LEVEL_NAMES), 4, n5, HORIZON, device=dev, level_id=lid))
    acts5 = [kb.random_actions(24, seed=1236 + r) for r, kb in enumerate(kbs)]
    k5 = [0]

    def step5():
        r, t_ = k5[0] % ring5, (k5[0] // ring5) % 24
        kbs[r].step(acts5[r][t_])
        k5[0] += 1

    t = timed(step5, 60, warm=8)
    b5 = n5 * (33 + 4 + 1)  # 16 + 16 + 4 action bytes + 1 reward/done byte + 1 level id byte
    out.append({"metric": "agent_steps_per_sec_cfg5_env", "value": n5 * 4 / t, "unit": "agent-steps/s",
                "config": "cfg-5 env half: 4 agents, nine-level mix (per-env level id), 2^20 envs per launch, plain Python loop",
                "us_per_step": t * 1e6,
                "roofline": {"bound": "hbm", "kernel": "step2_kernel<NA=4,NOBJ=4,EXTRAS=0,BITS=0,MULTI=1>", "bytes_per_unit": 38,
                             "achieved": b5 / t / 1e9, "peak": hbm_peak_gbs()[0], "unit": "GB/s",
                             "frac": b5 / t / 1e9 / hbm_peak_gbs()[0]}})
    del kbs, acts5
    # cfg-4: the device-resident Bayesian-Delegation loop (2-agent open-divider_salad, bd/bd): lower bounds +
    # exact Q through the planning-state memo + posterior + action selection + env step, 2^18 envs x 100
    # loop steps as SURVEY 8d states it, wall clock with a sync on both sides
    from gym_cooking_b200 import batched_agents
    n_loop, loop_steps = 1 << 18, 100
    loop = batched_agents.BatchedDelegation("open-divider_salad", n_loop, ("bd", "bd"), seed=1, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done_steps = loop.run(max_steps=loop_steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = loop.kb.stats().cpu().tolist()
    out.append({"metric": "bd_loop_agent_steps_per_sec", "value": loop.agent_steps / dt, "unit": "agent-steps/s",
                "config": "cfg-4: 2-agent open-divider_salad, bd/bd, 2^18 envs, horizon %d, from reset with a cold planner "
                          "memo; finished envs leave the working batch" % loop_steps,
                "posterior_updates_per_sec": loop.posterior_updates / dt, "seconds": dt, "loop_steps": done_steps,
                "delivered": st[1],
                "mean_steps_of_delivered": sum(t_ * c for t_, c in enumerate(st[5:5 + loop_steps])) / max(st[1], 1),
                "completed_subtasks": st[133], "planning_states_solved": loop.cache.solved_states,
                "planner_lookups": loop.cache.lookups})
    del loop
    # path B: cfg-3 (3 agents, full-divider_salad, 2^20 envs diversified by k = env % 41 random steps): lower
    # bounds of all 54 (subtask, agent set) pairs, then exact V* + Q[25] of every pair that is doable somewhere
    # in the batch, each distinct planning state solved once (subtask_q_unique)
    n = 1 << 20
    kb = gcb.KitchenBatch("full-divider_salad", 3, n, HORIZON, device=dev)
    acts = kb.random_actions(40, seed=1235)
    idx = torch.arange(n, device=dev) % 41
    for s_ in range(40):
        a = acts[s_].clone()
        a[idx <= s_] = 4
        kb.step(a)
    del acts
    ns = len(kb.subtasks[0])
    sets = [(i, None) for i in range(3)] + list(itertools.combinations(range(3), 2))
    pairs = [(s_, i, j) for s_ in range(ns) for (i, j) in sets]
    t = timed(lambda: gcb.lower_bound(kb, pairs), 5)
    lb_bytes = n * len(pairs) * 20  # 16 B state + 4 B bound per (env, pair)
    out.append({"metric": "lower_bounds_per_sec", "value": n * len(pairs) / t, "unit": "(env,pair)/s",
                "config": "cfg-3: 3-agent full-divider_salad, 2^20 envs x %d pairs" % len(pairs),
                "roofline": {"bound": "hbm", "kernel": "lower_bound_kernel", "bytes_per_unit": 20,
                             "achieved": lb_bytes / t / 1e9, "peak": hbm_peak_gbs()[0], "unit": "GB/s",
                             "frac": lb_bytes / t / 1e9 / hbm_peak_gbs()[0],
                             "note": "the state is read once per env, not per pair: compulsory traffic is 16 B + 4 B x 54"}})
    lb = gcb.lower_bound(kb, pairs)
    doable = [p for k, p in enumerate(pairs) if bool((lb[:, k] < 28).any())]
    del lb
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, q, status, n_unique = gcb.subtask_q_unique(kb, doable)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    hist = torch.bincount(status.flatten().long(), minlength=5).tolist()
    n_joint = sum(1 for p in doable if p[2] is not None)
    q_bytes = n * (n_joint * 120 + (len(doable) - n_joint) * 40)  # 40 B single / 120 B joint
    out.append({"metric": "subtask_values_per_sec", "value": n * len(doable) / t, "unit": "(env,pair)/s",
                "non_trivial_per_sec": (n * len(doable) - hist[2]) / t, "seconds": t,
                "config": "cfg-3: 2^20 envs x %d doable pairs (%d joint), exact V* + Q[25]; %d distinct planning states "
                          "solved once each (%.1fx)" % (len(doable), n_joint, n_unique, n / n_unique),
                "status_histogram": hist,
                "status_legend": "0 ok, 1 goal met at start, 2 unreachable (early exit), 3 search budget exceeded",
                "roofline": {"bound": "hbm", "kernel": "joint_tree_kernel + subtask_q_kernel", "achieved": q_bytes / t / 1e9,
                             "peak": hbm_peak_gbs()[0], "unit": "GB/s", "frac": q_bytes / t / 1e9 / hbm_peak_gbs()[0],
                             "note": "search-bound, not bandwidth-bound: the joint solver's open / closed sets live in a "
                                     "global-memory arena (profiles/r02_joint_tree_ncu.csv)"}})
    return out


GC_BENCH_MAX_PAIRS = 24


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cfg5-envs", type=int, default=256,
                    help="envs per level (x 9 levels) per GPU of the cfg-5 leg (4 agents, mixed models); 0 skips it")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the planner / posterior side metrics (used for short ncu passes)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 400:
            args.steps = 400  # bounded: ~30 ms per CPU step
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
