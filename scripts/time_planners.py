"""Device timing of the planner kernels (scratch helper; numbers quoted in DESIGN.md / profiles/)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb


def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def diversified(level, n_agents, n, seed=1235):
    kb = gcb.KitchenBatch(level, n_agents, n, 100)
    acts = kb.random_actions(40, seed=seed)
    idx = torch.arange(n, device=kb.device) % 41
    for s in range(40):
        a = acts[s].clone(); a[idx <= s] = 4
        kb.step(a)
    return kb


def posterior(n, H, P, A, E, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(1)
    probs = torch.rand((n, H), device="cuda", generator=g, dtype=dtype)
    hyp = torch.randint(0, P, (n, H, E), device="cuda", generator=g, dtype=torch.uint8)
    w = torch.randint(1, 3, (n, P), device="cuda", generator=g, dtype=torch.uint8)
    qd = torch.randn((n, P, A), device="cuda", generator=g, dtype=dtype)
    nv = torch.full((n, P), A, device="cuda", dtype=torch.uint8)
    ai = torch.randint(0, A, (n, P), device="cuda", generator=g, dtype=torch.uint8)
    t = timed(lambda: gcb.bd_posterior(probs, None, hyp, w, qd, nv, ai, 1.3), 10)
    es = 4 if dtype == torch.float32 else 8
    bytes_ = n * (2 * es * H + es * P * A + P + P + P + H * E)
    print("bd_posterior %s n=%d H=%d P=%d A=%d: %.1f us, %.3e updates/s, %.0f GB/s algorithmic (%.1f%% of 6453)" % (
        str(dtype)[6:], n, H, P, A, t * 1e6, n / t, bytes_ / t / 1e9, 100 * bytes_ / t / 6453.1e9))


if __name__ == "__main__":
    posterior(1 << 18, 8, 8, 5, 2)
    posterior(1 << 20, 8, 8, 5, 2)
    posterior(1 << 18, 36, 15, 25, 3)
    posterior(1 << 18, 84, 24, 25, 4)
    posterior(1 << 20, 8, 8, 5, 2, torch.float64)
    kb = diversified("full-divider_salad", 3, 1 << 20)
    ns = len(kb.subtasks[0])
    sets = [(i, None) for i in range(3)] + list(itertools.combinations(range(3), 2))
    pairs = [(s, i, j) for s in range(ns) for (i, j) in sets]
    t = timed(lambda: gcb.lower_bound(kb, pairs), 3)
    print("lower_bound cfg-3 2^20 envs x %d pairs: %.2f ms, %.3e (env,pair)/s" % (len(pairs), t * 1e3, (1 << 20) * len(pairs) / t))
    for level, na, n in (("full-divider_salad", 3, 1 << 16), ("open-divider_salad", 2, 1 << 16), ("partial-divider_tl", 2, 1 << 16)):
        kb = diversified(level, na, n)
        ns = len(kb.subtasks[0])
        pairs = [(s, i, None) for s in range(ns) for i in range(na)]
        out = {}
        def run():
            out["r"] = gcb.subtask_q(kb, pairs)
        t = timed(run, 2)
        v, q, status = out["r"]
        hist = torch.bincount(status.flatten().long(), minlength=5).tolist()
        print("subtask_q %s %d agents: %d envs x %d single pairs: %.1f ms, %.3e (env,pair)/s, status hist %s" % (
            level, na, n, len(pairs), t * 1e3, n * len(pairs) / t, hist))
