"""cfg-3 exact subtask values (single + joint) on diversified states (scratch timing helper)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
from time_planners import timed, diversified

for level, na, n in (("full-divider_salad", 3, 1 << 12), ("open-divider_salad", 2, 1 << 12), ("partial-divider_tl", 2, 1 << 12)):
    kb = diversified(level, na, n)
    ns = len(kb.subtasks[0])
    sets = list(itertools.combinations(range(na), 2))
    pairs = [(s, i, j) for s in range(ns) for (i, j) in sets]
    lb = gcb.lower_bound(kb, pairs)
    doable = [p for k, p in enumerate(pairs) if bool((lb[:, k] < 28).any())][:24]
    res = {}
    t = timed(lambda: res.update(r=gcb.subtask_q(kb, doable)), 1)
    st = res["r"][2]
    print("%s: %d envs x %d joint pairs: %.2f s, %.3e (env,pair)/s, status %s" % (
        level, n, len(doable), t, n * len(doable) / t, torch.bincount(st.flatten().long(), minlength=5).tolist()), flush=True)
