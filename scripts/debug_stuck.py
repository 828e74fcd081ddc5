import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb
from gym_cooking_b200 import batched_agents
level = sys.argv[1] if len(sys.argv) > 1 else "open-divider_salad"
loop = batched_agents.BatchedDelegation(level, 2048, ("bd", "bd"), seed=1)
for s in range(60): loop.step()
und = (~loop.kb.done).nonzero()[:, 0]
print("undone", len(und), "subtasks", [str(s) for s in loop.subtasks])
c = collections.Counter()
for e in und[:400].tolist():
    d = gcb.decode_state(loop.kb.state[e].tolist(), 2)
    objs = tuple(sorted((m, h != 0) for m, x, y, h in d["objects"]))
    c[(objs, bin(int(loop.incomplete[e, 0])), bin(int(loop.incomplete[e, 1])), int(loop.alive[0][e].sum()), int(loop.alive[1][e].sum()))] += 1
for k, v in c.most_common(12): print(v, k)
e = und[0].item()
print(gcb.decode_state(loop.kb.state[e].tolist(), 2), loop.cur_sub[e].tolist())
for i in range(2):
    T = loop.tables[i]
    print("agent", i, [(T.keys[h], round(float(loop.probs[i][e, h]), 4)) for h in range(T.H) if bool(loop.alive[i][e, h])])
print("---- livelock example")
for e in und.tolist():
    d = gcb.decode_state(loop.kb.state[e].tolist(), 2)
    if len(d["objects"]) == 4:
        break
for s in range(8):
    d = gcb.decode_state(loop.kb.state[e].tolist(), 2)
    print(d["agents"], d["objects"])
    loop.step()
    print("   subs", loop.cur_sub[e].tolist(), loop.cur_joint[e].tolist(), "actions", loop.last_actions[e].tolist(), "executed", loop.executed[e].tolist())
    for i in range(2):
        T = loop.tables[i]
        print("   agent", i, [(T.keys[h], round(float(loop.probs[i][e, h]), 3)) for h in range(T.H) if bool(loop.alive[i][e, h])])
