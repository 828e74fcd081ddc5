"""cfg-3: full Bayesian-Delegation loop over a batch (scratch timing helper)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_cooking_b200 import batched_agents

level = sys.argv[1] if len(sys.argv) > 1 else "open-divider_salad"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
max_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
loop = batched_agents.BatchedDelegation(level, n, ("bd", "bd"), seed=1)
torch.cuda.synchronize(); t0 = time.time()
steps = loop.run(max_steps=max_steps)
n_work = loop.wkb.num_envs
torch.cuda.synchronize(); dt = time.time() - t0
print("working batch at the end: %d envs" % n_work)
stats = loop.kb.stats().cpu().numpy()
t = ((loop.kb.state[:, 0].to(torch.int64) >> 24) & 127).float()
print("%s n=%d: %d loop steps in %.2f s; delivered %d (%.1f%%), mean t %.1f; %d posterior updates (%.3e/s), %.3e agent-steps/s, %d planning states solved for %d lookups" % (
    level, n, steps, dt, int(stats[1]), 100.0 * int(stats[1]) / n, float(t.mean()), loop.posterior_updates, loop.posterior_updates / dt,
    float(t.sum()) * 2 / dt, loop.cache.solved_states, loop.cache.lookups))
