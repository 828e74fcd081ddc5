import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
N = 1 << 20
h = torch.randint(0, 5, (N, 2), dtype=torch.uint8).pin_memory()
def wall(fn, iters=300):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / iters * 1e6
ns = argparse.Namespace(level="partial-divider_tl", num_agents=2, max_num_timesteps=0, max_num_subtasks=14, seed=1, model1=None, model2=None, model3=None, model4=None)
env = gcb.OvercookedEnvironment(ns, num_envs=N, track_collisions=False); env.reset()
for piece in (4096, 2048, 1024, 512, 256, 128):
    env.H2D_PIECE_BYTES = piece << 10
    print("facade step, H2D pieces of %4d KB: %.1f us" % (piece, wall(lambda: env.step(h))))
