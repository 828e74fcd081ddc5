"""Single-agent exact subtask values (gc_subtask_q, IDA*) on diversified states: seconds per launch and status counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
from time_planners import timed, diversified

for level, na, n in (("partial-divider_tl", 4, 1 << 12), ("full-divider_salad", 3, 1 << 12), ("open-divider_salad", 2, 1 << 14),
                     ("partial-divider_salad", 4, 1 << 12)):
    kb = diversified(level, na, n)
    ns = len(kb.subtasks[0])
    pairs = [(s, i, None, lvl) for s in range(ns) for i in range(na) for lvl in ((False, True) if na > 1 else (False,))]
    res = {}
    t = timed(lambda: res.update(r=gcb.subtask_q(kb, pairs)), 2)
    v, q, st = res["r"]
    print("%s (%d agents): %d envs x %d single pairs: %.3f s, %.3e (env,pair)/s, status %s, checksum %.6f" % (
        level, na, n, len(pairs), t, n * len(pairs) / t, torch.bincount(st.flatten().long(), minlength=5).tolist(),
        float(torch.nan_to_num(q, nan=0.0, posinf=1000.0).double().sum())), flush=True)
