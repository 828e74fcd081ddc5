"""Quick device timing of the env-step kernel (scratch helper; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb

def time_step(level, n_agents, n, ring=16, iters=400):
    kbs = [gcb.KitchenBatch(level, n_agents, n, 100) for _ in range(ring)]
    acts = [kb.random_actions(100, seed=7 + i) for i, kb in enumerate(kbs)]
    for w in range(3):
        for r in range(ring):
            kbs[r].step(acts[r][w])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(iters):
        r = it % ring
        kbs[r].step(acts[r][3 + it // ring])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bytes_ = n * (33 + n_agents)
    print("%s n_agents=%d n=%d: %.2f us/step, %.3e agent-steps/s, %.1f GB/s algorithmic (%.1f%% of 6453)" % (
        level, n_agents, n, ms * 1e3, n * n_agents / (ms * 1e-3), bytes_ / (ms * 1e-3) / 1e9,
        100 * bytes_ / (ms * 1e-3) / 6453.1e9))

if __name__ == "__main__":
    time_step("partial-divider_tl", 2, 1 << 20)
    time_step("partial-divider_tl", 2, 1 << 24, ring=2, iters=60)
    time_step("full-divider_salad", 3, 1 << 20)
    time_step("open-divider_salad", 4, 1 << 20)
    kb = gcb.KitchenBatch("partial-divider_tl", 2, 1 << 20, 100)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kb.rollout(100, seed=1); kb.reset(); torch.cuda.synchronize()
    e0.record(); kb.rollout(100, seed=2); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("rollout 2^20 envs x 100 steps: %.3f ms, %.3e agent-steps/s" % (ms, (1 << 20) * 100 * 2 / (ms * 1e-3)))
