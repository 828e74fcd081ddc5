"""cfg-3 (3-agent full-divider_salad, diversified states): which planner kernel the time goes to (torch profiler).
usage: python scripts/prof_cfg3.py [log2_envs]"""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
from time_planners import diversified
from torch.profiler import ProfilerActivity, profile

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 18)
kb = diversified("full-divider_salad", 3, n)
ns = len(kb.subtasks[0])
sets = [(i, None) for i in range(3)] + list(itertools.combinations(range(3), 2))
pairs = [(s, i, j) for s in range(ns) for (i, j) in sets]
lb = gcb.lower_bound(kb, pairs)
doable = [p for k, p in enumerate(pairs) if bool((lb[:, k] < 28).any())]
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    v, q, st, nu = gcb.subtask_q_unique(kb, doable)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print("%d envs, %d doable pairs, %d distinct states: %.2f s, status %s" % (
    n, len(doable), nu, dt, torch.bincount(st.flatten().long(), minlength=5).tolist()))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
