"""Where the time of one host-driven step goes (2^20 envs, 2 agents): per-call microseconds of each layer
and of the individual stream operations."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb

N, NA = 1 << 20, 2
dev = torch.device("cuda", 0)
ns = argparse.Namespace(level="partial-divider_tl", num_agents=NA, max_num_timesteps=100, max_num_subtasks=14, seed=1,
                        model1=None, model2=None, model3=None, model4=None)
env = gcb.OvercookedEnvironment(ns, num_envs=N, device=dev, track_collisions=False)
env.reset()
kb = env._kb
acts = kb.random_actions(8, seed=3)
byte_h = [acts[s].cpu().pin_memory() for s in range(8)]
joint_h = [(acts[s][:, 0] * 5 + acts[s][:, 1]).to(torch.uint8).cpu().pin_memory() for s in range(8)]
joint_d = [j.to(dev) for j in joint_h]
bits = torch.zeros(((N + 31) // 32, 2), dtype=torch.int32).pin_memory()
bits_d = torch.zeros(((N + 31) // 32, 2), dtype=torch.int32, device=dev)
stage = torch.empty(N, dtype=torch.uint8, device=dev)


def timeit(name, fn, reps=300, warm=300):
    for k in range(warm):
        if k % 90 == 0:
            kb.reset()
        fn(k)
    torch.cuda.synchronize()
    kb.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps):
        if k % 90 == 89:
            kb.reset()
        fn(k)
    torch.cuda.synchronize()
    us = (time.perf_counter() - t0) / reps * 1e6
    print("%-58s %7.1f us per step  (%.2e agent-steps/s)" % (name, us, N * NA / us * 1e6))


timeit("env.step(joint host)", lambda k: env.step(joint_h[k % 8]))
timeit("env.step(bytes host)", lambda k: env.step(byte_h[k % 8]))
timeit("kb.step_host_bits(joint host)", lambda k: kb.step_host_bits(joint_h[k % 8], bits))
timeit("kb.step_host_bits(bytes host)", lambda k: kb.step_host_bits(byte_h[k % 8], bits))
timeit("kb.step(joint device) + sync", lambda k: (kb.step(joint_d[k % 8]), torch.cuda.synchronize()))
timeit("kb.step(joint device), no sync", lambda k: kb.step(joint_d[k % 8]))
timeit("H2D 1 MB + sync", lambda k: (stage.copy_(joint_h[k % 8], non_blocking=True), torch.cuda.synchronize()))
timeit("D2H 256 KB + sync", lambda k: (bits.copy_(bits_d, non_blocking=True), torch.cuda.synchronize()))
timeit("H2D 1 MB + step + D2H 256 KB + sync (torch ops)",
       lambda k: (stage.copy_(joint_h[k % 8], non_blocking=True), kb.step(stage), bits.copy_(bits_d, non_blocking=True),
                  torch.cuda.synchronize()))
timeit("sync only", lambda k: torch.cuda.synchronize())
