#!/bin/bash
# cfg-4 loop time vs the joint solver's wide/narrow switch (GC_JOINT_WIDE_PROBLEMS: problem count up to which a
# launch uses 2 CTAs of 512 threads per SM instead of 16 of 64)
cd "$(dirname "$0")/.."
for w in ${GC_WIDE_LIST:-32768 4096 1024 128 0}; do
  echo "== GC_JOINT_WIDE_PROBLEMS=$w"
  GC_JOINT_WIDE_PROBLEMS=$w python - <<'PY'
import time, torch, sys, os
sys.path.insert(0, os.getcwd())
from gym_cooking_b200 import batched_agents
loop = batched_agents.BatchedDelegation("open-divider_salad", 1 << 18, ("bd", "bd"), seed=1)
torch.cuda.synchronize(); t0 = time.perf_counter()
steps = loop.run(max_steps=100)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
st = loop.kb.stats().cpu().tolist()
print("%d loop steps %.2f s, %.3g agent-steps/s, delivered %d, states solved %d" % (steps, dt, loop.agent_steps / dt, st[1], loop.cache.solved_states))
PY
done
