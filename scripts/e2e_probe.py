"""Where does an end-to-end batched step spend its time? (scratch helper)"""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb

N = 1 << 20
dev = torch.device("cuda")
h = torch.randint(0, 5, (N, 2), dtype=torch.uint8).pin_memory()
d = torch.empty((N, 2), dtype=torch.uint8, device=dev)
rd = torch.zeros(N, dtype=torch.uint8, device=dev)
hrd = torch.empty(N, dtype=torch.uint8).pin_memory()

def wall(fn, iters=200):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / iters * 1e6

print("H2D 2MB + sync: %.1f us" % wall(lambda: (d.copy_(h, non_blocking=True), torch.cuda.current_stream().synchronize())))
print("D2H 1MB + sync: %.1f us" % wall(lambda: (hrd.copy_(rd, non_blocking=True), torch.cuda.current_stream().synchronize())))
print("H2D 1MB + sync: %.1f us" % wall(lambda: (d[:N // 2].copy_(h[:N // 2], non_blocking=True), torch.cuda.current_stream().synchronize())))
print("D2H 256KB + sync: %.1f us" % wall(lambda: (hrd[:N // 4].copy_(rd[:N // 4], non_blocking=True), torch.cuda.current_stream().synchronize())))
kb = gcb.KitchenBatch("partial-divider_tl", 2, N, 0)
print("kernel + sync: %.1f us" % wall(lambda: (kb.step(d), torch.cuda.current_stream().synchronize())))
print("H2D+kernel+D2H + sync: %.1f us" % wall(lambda: (d.copy_(h, non_blocking=True), kb.step(d), hrd.copy_(kb.reward_done, non_blocking=True), torch.cuda.current_stream().synchronize())))
ns = argparse.Namespace(level="partial-divider_tl", num_agents=2, max_num_timesteps=0, max_num_subtasks=14, seed=1, model1=None, model2=None, model3=None, model4=None)
env = gcb.OvercookedEnvironment(ns, num_envs=N, track_collisions=False); env.reset()
for c in (1, 2, 4):
    env.PIPELINE_CHUNKS = c; env._streams = None
    print("facade step, %d chunk(s): %.1f us" % (c, wall(lambda: env.step(h))))
env.PIPELINE_MIN_ENVS = 1 << 30
print("facade step, unpipelined: %.1f us" % wall(lambda: env.step(h)))
