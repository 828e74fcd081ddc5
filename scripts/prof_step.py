"""Minimal workload for ncu: a few env-step launches of cfg-2 (2 agents, partial-divider_tl, 2^20 envs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n_agents = int(sys.argv[2]) if len(sys.argv) > 2 else 2
level = sys.argv[3] if len(sys.argv) > 3 else "partial-divider_tl"
kb = gcb.KitchenBatch(level, n_agents, n, 100)
acts = kb.random_actions(40, seed=1234)
for s in range(40):
    kb.step(acts[s])
torch.cuda.synchronize()
print("ok", int(kb.state[:, 0].sum()))
