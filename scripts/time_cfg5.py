"""cfg-5: 4 agents, mixed bd/up/dc/fb/greedy, all nine levels - wall time of the delegation loop per level.
usage: python scripts/time_cfg5.py [envs_per_level] [max_steps] [levels,comma,separated]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gym_cooking_b200 import batched_agents  # noqa: E402

MODELS = ("bd", "up", "dc", "fb", "greedy")
LEVELS = ["%s-divider_%s" % (d, r) for r in ("tomato", "tl", "salad") for d in ("open", "partial", "full")]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 12
    max_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    levels = sys.argv[3].split(",") if len(sys.argv) > 3 else LEVELS
    prof = os.environ.get("GC_PROF")
    for k, level in enumerate(levels):
        models = tuple(MODELS[(k + j) % 5] for j in range(4))
        t0 = time.perf_counter()
        loop = batched_agents.BatchedDelegation(level, n, models, seed=1 + k)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        widths = []
        if prof:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
                steps = loop.run(max_steps=max_steps)
                torch.cuda.synchronize()
            print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25))
        else:
            steps = loop.run(max_steps=max_steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t1
        st = loop.kb.stats().cpu().tolist()
        print("%-24s %s H=%s: setup %.1f s, %d loop steps %.2f s, %.3g agent-steps/s, delivered %d/%d, states solved %d "
              "for %d lookups, list widths %s, mem %.1f GB" % (
                  level, "/".join(models), [T.H for T in loop.tables], t1 - t0, steps, dt, loop.agent_steps / dt, st[1], n,
                  loop.cache.solved_states, loop.cache.lookups,
                  [int(a.shape[1]) for a in loop.alive], torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
        if os.environ.get("GC_STATUS"):
            stt = loop.cache.status  # [states][pairs]
            kinds = {}
            for k_, pr in enumerate(loop.cache.pairs):
                kinds.setdefault(("joint" if pr[2] is not None else "single", "L1" if pr[3] else "L0"), []).append(k_)
            for key, cols in kinds.items():
                h = torch.bincount(stt[:, cols].flatten().long(), minlength=5).tolist()
                print("   %s %s: status histogram %s (0 ok, 1 goal met, 2 unreachable, 3 budget, 4 unsupported)" % (key[0], key[1], h))
        del loop
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
