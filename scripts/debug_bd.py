"""Step the facade loop and the batched loop side by side (deterministic ties) and print the first divergence."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gym_cooking_b200 as gcb
from gym_cooking_b200 import batched_agents, delegation_planner, main as gmain, navigation_planner
from gym_cooking_b200.utils.agent import RealAgent

level, models = sys.argv[1], tuple(sys.argv[2].split(","))
loop = batched_agents.BatchedDelegation(level, 4, models, deterministic=True)
subtasks = loop.subtasks
def get_max(self):
    if not self.probs: return None
    best = max(self.probs.values())
    c = [a for a, p in self.probs.items() if p >= best - 1e-12]
    return min(c, key=lambda a: batched_agents.alloc_key(a, subtasks))
delegation_planner.SubtaskAllocDistribution.get_max = get_max
navigation_planner.argmin = lambda v: int(np.argmin(np.asarray(v, dtype=np.float64)))
np.random.choice = lambda n, p=None: n - 1
ms = list(models) + [None] * (4 - len(models))
arglist = argparse.Namespace(level=level, num_agents=len(models), max_num_timesteps=100, max_num_subtasks=14, seed=1, beta=1.3, alpha=0.01, tau=2, cap=75, main_cap=100, play=False, record=False, with_image_obs=False, model1=ms[0], model2=ms[1], model3=ms[2], model4=ms[3])
env = gcb.make(arglist=arglist); obs = env.reset()
agents = gmain.initialize_agents(arglist, env)
for step in range(60):
    if env.done(): break
    ad = {a.name: a.select_action(obs=obs) for a in agents}
    loop.step()
    want = [gcb.ACTION_INDEX[tuple(ad[a.name])] for a in agents]
    got = loop.last_actions[0].tolist()
    print("step", step, "facade", want, [(str(a.subtask), a.subtask_agent_names) for a in agents], "batched", got, loop.cur_sub[0].tolist(), loop.cur_joint[0].tolist())
    if want != got:
        for i, a in enumerate(agents):
            T = loop.tables[i]
            fp = {batched_agents.alloc_key(k, subtasks): p for k, p in a.delegator.probs.probs.items()}
            bp = {T.keys[h]: float(loop.probs[i][0, h]) for h in range(T.H) if bool(loop.alive[i][0, h])}
            print(" agent", i, "incomplete facade", [str(s) for s in a.incomplete_subtasks], "batched", bin(int(loop.incomplete[0, i])))
            for k in sorted(set(fp) | set(bp)):
                print("   %-40s facade %-22s batched %s" % (k, fp.get(k), bp.get(k)))
        break
    obs, _, _, _ = env.step(action_dict=ad)
    for a in agents: a.refresh_subtasks(world=env.world)
