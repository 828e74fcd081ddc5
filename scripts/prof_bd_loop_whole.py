"""torch profiler over a WHOLE cfg-4 run (2^18 envs x 100 loop steps, bd/bd): share of the device time per kernel."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from gym_cooking_b200 import batched_agents
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
loop = batched_agents.BatchedDelegation("open-divider_salad", n, ("bd", "bd"), seed=1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.time()
    steps = loop.run(max_steps=100)
    torch.cuda.synchronize()
    dt = time.time() - t0
print("%d envs, %d loop steps: %.2f s wall under the profiler" % (n, steps, dt))
ka = prof.key_averages()
tot = sum(e.self_device_time_total for e in ka)
ours = ("joint_", "subtask_q", "lower_bound", "bd_", "step2", "stats", "reset", "pack_")
mine = sum(e.self_device_time_total for e in ka if any(k in e.key for k in ours))
print("device time %.2f s, of which repo kernels %.2f s = %.1f %%" % (tot / 1e6, mine / 1e6, 100 * mine / tot))
print(ka.table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=56))
