"""Does binding the host thread to the GPU's NUMA node change the end-to-end step rate? (scratch probe)
Blocks of 400 gym-style steps through OvercookedEnvironment.step with pinned host actions, before and
after nvmlDeviceSetCpuAffinity; pinned buffers are re-allocated after the bind."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb

N = 1 << 20
ns = argparse.Namespace(level="partial-divider_tl", num_agents=2, max_num_timesteps=100, max_num_subtasks=14,
                        seed=1, model1=None, model2=None, model3=None, model4=None)


def blocks(tag, n_blocks=5, steps=400):
    env = gcb.OvercookedEnvironment(ns, num_envs=N, track_collisions=False)
    env.reset()
    kb = gcb.KitchenBatch("partial-divider_tl", 2, N, 100)
    acts = kb.random_actions(8, seed=3)
    host = [acts[s].cpu().pin_memory() for s in range(8)]
    for s in range(3):
        env.step(host[s])
    out = []
    for b in range(n_blocks):
        env.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(steps):
            if s and s % 100 == 0:
                env.reset()
            env.step(host[s % 8])
        torch.cuda.synchronize()
        out.append(steps * N * 2 / (time.perf_counter() - t0))
    print(tag, " ".join("%.3e" % v for v in out), "affinity", sorted(os.sched_getaffinity(0)))


os.system("nvidia-smi topo -m 2>&1 | head -12; lscpu | grep -i 'numa\\|^CPU(s)\\|Model name' ")
blocks("default ")
import pynvml
pynvml.nvmlInit()
uuid = str(torch.cuda.get_device_properties(0).uuid)
h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
try:
    pynvml.nvmlDeviceSetCpuAffinity(h)
    print("bound to the GPU's ideal CPUs")
except Exception as exc:
    print("bind failed:", exc)
blocks("bound   ")
