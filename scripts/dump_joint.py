"""Dump V/Q/status of the joint solver on diversified states (run once per solver variant, then compare)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import gym_cooking_b200 as gcb
from time_planners import diversified
out = {}
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 11
for level, na, n in (("full-divider_salad", 3, N), ("open-divider_salad", 2, N), ("partial-divider_tl", 2, N), ("open-divider_tl", 2, N)):
    kb = diversified(level, na, n)
    ns = len(kb.subtasks[0])
    sets = list(itertools.combinations(range(na), 2))
    pairs = [(s, i, j) for s in range(ns) for (i, j) in sets]
    lb = gcb.lower_bound(kb, pairs)
    doable = [p for k, p in enumerate(pairs) if bool((lb[:, k] < 28).any())][:24]
    v, q, st = gcb.subtask_q(kb, doable)
    out[level + "_v"], out[level + "_q"], out[level + "_s"] = v.cpu().numpy(), q.cpu().numpy(), st.cpu().numpy()
np.savez(sys.argv[1], **out)
