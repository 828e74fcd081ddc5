"""cfg-3 (3 agents, full-divider_salad): distinct planning states and planner time as the batch grows."""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb


def diversified(n, seed=1235):
    kb = gcb.KitchenBatch("full-divider_salad", 3, n, 100)
    acts = kb.random_actions(40, seed=seed)
    idx = torch.arange(n, device=kb.device) % 41
    for s in range(40):
        a = acts[s].clone()
        a[idx <= s] = 4
        kb.step(a)
    return kb


for logn in (12, 16, 18, 20) if not os.environ.get("GC_JOINT_CTAS_PER_SM") else (16, 18):
    n = 1 << logn
    kb = diversified(n)
    ns = len(kb.subtasks[0])
    sets = [(i, None) for i in range(3)] + list(itertools.combinations(range(3), 2))
    pairs = [(s, i, j) for s in range(ns) for (i, j) in sets]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lb = gcb.lower_bound(kb, pairs)
    torch.cuda.synchronize()
    t_lb = time.perf_counter() - t0
    doable = [p for k, p in enumerate(pairs) if bool((lb[:, k] < 28).any())]
    t0 = time.perf_counter()
    v, q, status, U = gcb.subtask_q_unique(kb, doable)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    hist = torch.bincount(status.flatten().long(), minlength=5).tolist()
    nontrivial = n * len(doable) - hist[2]
    print("n=2^%d: %d pairs (%d doable, %d joint), lower bounds %.3f s; %d distinct planning states (%.1fx), "
          "subtask_q_unique %.2f s -> %.3g (env,pair)/s, %.3g non-trivial/s; status %s"
          % (logn, len(pairs), len(doable), sum(1 for p in doable if p[2] is not None), t_lb, U, n / U, dt,
             n * len(doable) / dt, nontrivial / dt, hist), flush=True)
    if logn == 12:  # brute force for the ratio, and equality with the memoised answers
        t0 = time.perf_counter()
        v2, q2, s2 = gcb.subtask_q(kb, doable)
        torch.cuda.synchronize()
        print("   brute force: %.2f s; equal: %s" % (time.perf_counter() - t0, bool(torch.equal(status, s2) and torch.equal(
            torch.nan_to_num(v, posinf=1e9), torch.nan_to_num(v2, posinf=1e9)))))
    del kb
