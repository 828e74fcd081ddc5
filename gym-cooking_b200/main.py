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


def main_loop(arglist, max_steps=None, bag_directory=None):
    """Runs one episode; returns (env, real_agents, history of action dicts).  With `bag_directory`
    the episode is also written as the reference's `Bag` pickle (main.py:94-117)."""
    env = make("gym_cooking:overcookedEnv-v0", arglist=arglist)
    obs = env.reset()
    agents = initialize_agents(arglist, env)
    bag = None
    if bag_directory is not None:
        from .misc.metrics.metrics_bag import Bag
        bag = Bag(arglist=arglist, filename=env.filename, directory=bag_directory)
        bag.set_recipe(recipe_subtasks=env.all_subtasks)
    history = []
    while not env.done():
        action_dict = {a.name: a.select_action(obs=obs) for a in agents}
        obs, reward, done, info = env.step(action_dict=action_dict)
        for a in agents:
            a.refresh_subtasks(world=env.world)
        if bag is not None:
            bag.add_status(cur_time=info["t"], real_agents=agents)
        history.append(action_dict)
        if max_steps is not None and len(history) >= max_steps:
            break
    if bag is not None:
        bag.set_collisions(collisions=env.collisions)
        env.bag_path = bag.set_termination(termination_info=env.termination_info, successful=env.successful)
    return env, agents, history
