"""`main_loop` of the reference (main.py:85-117) on the B200-native environment and planners."""
import random

import numpy as np

from . import make
from .utils.agent import COLORS, RealAgent


def fix_seed(seed):  # main.py:53-55
    np.random.seed(seed)
    random.seed(seed)


def initialize_agents(arglist, env):  # main.py:57-83
    return [RealAgent(arglist=arglist, name="agent-%d" % (i + 1), id_color=COLORS[i], recipes=env.recipes)
            for i in range(arglist.num_agents)]


def main_loop(arglist, max_steps=None):
    """Runs one episode; returns (env, real_agents, history of action dicts)."""
    env = make("gym_cooking:overcookedEnv-v0", arglist=arglist)
    obs = env.reset()
    agents = initialize_agents(arglist, env)
    history = []
    while not env.done():
        action_dict = {a.name: a.select_action(obs=obs) for a in agents}
        obs, reward, done, info = env.step(action_dict=action_dict)
        for a in agents:
            a.refresh_subtasks(world=env.world)
        history.append(action_dict)
        if max_steps is not None and len(history) >= max_steps:
            break
    return env, agents, history
