import sys, os, itertools, collections
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
from time_planners import diversified
level, na, n = "open-divider_salad", 2, 1 << 12
kb = diversified(level, na, n)
ns = len(kb.subtasks[0])
pairs = [(s, 0, 1) for s in range(ns)]
lb = gcb.lower_bound(kb, pairs)
doable = [p for k, p in enumerate(pairs) if bool((lb[:, k] < 28).any())][:24]
v, q, st = gcb.subtask_q(kb, doable)
print([str(kb.subtasks[0][p[0]]) for p in doable])
print("status by pair:", [(str(kb.subtasks[0][p[0]]), torch.bincount(st[:, k].long(), minlength=4).tolist()) for k, p in enumerate(doable)])
c = collections.Counter()
for e, k in (st == 3).nonzero()[:4000].tolist():
    d = gcb.decode_state(kb.state[e].tolist(), na)
    objs = tuple(sorted((m, h != 0) for m, x, y, h in d["objects"]))
    c[(str(kb.subtasks[0][doable[k][0]]), objs)] += 1
for key, cnt in c.most_common(12): print(cnt, key)
e, k = (st == 3).nonzero()[0].tolist()
print(gcb.decode_state(kb.state[e].tolist(), na), str(kb.subtasks[0][doable[k][0]]), "lb", float(lb[e, pairs.index(doable[k])]), "v", float(v[e, k]), q[e, k].tolist())
