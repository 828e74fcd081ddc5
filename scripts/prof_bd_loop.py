"""torch profiler over steady-state steps of the batched delegation loop (scratch)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from gym_cooking_b200 import batched_agents
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
loop = batched_agents.BatchedDelegation("open-divider_salad", n, ("bd", "bd"), seed=1)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12): loop.step()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(3): loop.step()
torch.cuda.synchronize(); print("3 steps %.3f s" % (time.time() - t0))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): loop.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
