"""Small launches of every planner / delegation kernel, for ncu (profiles/r02_*): lower_bound_kernel,
subtask_q_kernel, joint_tree_kernel (+ joint_q_kernel), bd_rows_kernel, bd_posterior_tiles_kernel,
step2_kernel<4,4,EXTRAS=0,MULTI=1>."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
from gym_cooking_b200 import batched_agents

n = 1 << 12
kb = gcb.KitchenBatch("full-divider_salad", 3, n, 100)
acts = kb.random_actions(40, seed=1235)
idx = torch.arange(n, device=kb.device) % 41
for s in range(40):
    a = acts[s].clone()
    a[idx <= s] = 4
    kb.step(a)
ns = len(kb.subtasks[0])
sets = [(i, None) for i in range(3)] + list(itertools.combinations(range(3), 2))
pairs = [(s, i, j) for s in range(ns) for (i, j) in sets]
big = gcb.KitchenBatch("full-divider_salad", 3, 1 << 18, 100)
big.rollout(30, seed=5)
g = torch.Generator().manual_seed(5)
m = 1 << 20
lid = torch.randint(0, 9, (m,), generator=g, dtype=torch.uint8)
mk = gcb.KitchenBatch(list(gcb.levels.LEVEL_NAMES), 4, m, 100, level_id=lid)
ma = mk.random_actions(6, seed=1236)
loop = batched_agents.BatchedDelegation("open-divider_salad", 1 << 14, ("bd", "bd"), seed=1)
for _ in range(4):
    loop.step()
gcb.lower_bound(big, pairs)
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: everything above is set-up
for _ in range(1):
    gcb.lower_bound(big, pairs)
lb = gcb.lower_bound(kb, pairs)
doable = [p for k, p in enumerate(pairs) if bool((lb[:, k] < 28).any())]
for _ in range(1):
    gcb.subtask_q(kb, doable)
# cfg-5 style env step
for s in range(2):
    mk.step(ma[s])
# delegation loop kernels
loop.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
