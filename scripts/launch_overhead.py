"""Host cost of one KitchenBatch.step call (tiny batch, so the kernel is negligible) and the step rate at 2^20 envs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
kb = gcb.KitchenBatch("partial-divider_tl", 2, 64, 100)
a = kb.random_actions(1)[0]
for _ in range(100): kb.step(a)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20000): kb.step(a)
t1 = time.perf_counter(); torch.cuda.synchronize()
print("host time per step call: %.2f us" % ((t1 - t0) / 20000 * 1e6))
from quick_time import time_step
time_step("partial-divider_tl", 2, 1 << 20)
time_step("partial-divider_tl", 2, 1 << 20)
