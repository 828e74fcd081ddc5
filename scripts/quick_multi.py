"""4-agent step timing: single level (table form vs generic form) and the nine-level mix (scratch helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
from quick_time import time_step

def time_multi(n_agents, n, ring=8, iters=200):
    g = torch.Generator().manual_seed(5)
    kbs = []
    for r in range(ring):
        level_id = torch.randint(0, 9, (n,), generator=g, dtype=torch.uint8)
        kbs.append(gcb.KitchenBatch(list(gcb.levels.LEVEL_NAMES), n_agents, n, 100, level_id=level_id))
    acts = [kb.random_actions(60, seed=7 + i) for i, kb in enumerate(kbs)]
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
    bytes_ = n * (34 + n_agents)
    print("nine levels n_agents=%d n=%d: %.2f us/step, %.3e agent-steps/s, %.1f GB/s algorithmic (%.1f%% of 6453)" % (
        n_agents, n, ms * 1e3, n * n_agents / (ms * 1e-3), bytes_ / (ms * 1e-3) / 1e9, 100 * bytes_ / (ms * 1e-3) / 6453.1e9))

time_step("open-divider_salad", 4, 1 << 20)
time_multi(4, 1 << 20)
time_multi(2, 1 << 20)
