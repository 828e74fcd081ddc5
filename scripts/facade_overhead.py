"""Host cost of one batched facade step (tiny batch: copies and kernel negligible) vs the bare library call."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
ns = argparse.Namespace(level="partial-divider_tl", num_agents=2, max_num_timesteps=100, max_num_subtasks=14, seed=1,
                        model1=None, model2=None, model3=None, model4=None)
n = 64
env = gcb.OvercookedEnvironment(ns, num_envs=n, track_collisions=False)
env.reset()
a = torch.zeros((n, 2), dtype=torch.uint8).pin_memory()
for _ in range(200): env.step(a)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5000): env.step(a)
t1 = time.perf_counter()
print("facade step (host in, host out): %.1f us" % ((t1 - t0) / 5000 * 1e6))
kb = env._kb
t0 = time.perf_counter()
for _ in range(5000): kb.step_host(a, env._dev_actions, env._pinned_rd)
t1 = time.perf_counter()
print("KitchenBatch.step_host alone: %.1f us" % ((t1 - t0) / 5000 * 1e6))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): env.step(a)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
