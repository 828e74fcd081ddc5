"""SimAgent view (utils/agent.py:371-423 of the reference): env-side body of one agent."""
from collections import namedtuple

AgentRepr = namedtuple("AgentRepr", "name location holding")
COLORS = ["blue", "magenta", "yellow", "green"]


class SimAgent:
    def __init__(self, name, id_color, location):
        self.name = name
        self.color = id_color
        self.location = location
        self.holding = None
        self.action = (0, 0)
        self.has_delivered = False

    def get_repr(self):
        return AgentRepr(name=self.name, location=self.location, holding=self.get_holding())

    def get_holding(self):
        return "None" if self.holding is None else self.holding.full_name

    def __str__(self):  # agent.py:382-383 (without the terminal colour codes)
        return self.name[-1]

    def print_status(self):  # agent.py:401-406
        print("{} currently at {}, action {}, holding {}".format(self.name, self.location, self.action, self.get_holding()))

    def __repr__(self):
        return "SimAgent(%s @%s holding %s)" % (self.name, self.location, self.get_holding())


class RealAgent:
    """Host mirror of the reference's RealAgent (utils/agent.py:28-368): recipe subtasks ->
    Bayesian Delegation -> navigation planner, one action per call.  All planning numbers come
    from the kernels through the E2E_BRTDP / BayesianDelegator facades; this class is the
    control flow around them (reset-vs-update :176-203, plan :218-281, completion :286-368)."""

    def __init__(self, arglist, name, id_color, recipes):
        from .. import navigation_planner
        self.arglist, self.name, self.color, self.recipes = arglist, name, id_color, recipes
        self.reset_subtasks()
        self.new_subtask, self.new_subtask_agent_names = None, []
        self.incomplete_subtasks = []
        self.is_subtask_complete = lambda w: False
        self.beta = getattr(arglist, "beta", 1.3)
        self.none_action_prob = 0.5
        self.model_type = getattr(arglist, "model%s" % name[-1])
        self.priors = "uniform" if self.model_type == "up" else "spatial"
        self.planner = navigation_planner.E2E_BRTDP(alpha=getattr(arglist, "alpha", 0.01), tau=getattr(arglist, "tau", 2),
                                                    cap=getattr(arglist, "cap", 75), main_cap=getattr(arglist, "main_cap", 100))
        self.location, self.holding, self.action = None, None, (0, 0)

    def get_holding(self):
        return "None" if self.holding is None else self.holding.full_name

    def reset_subtasks(self):
        self.subtask, self.subtask_agent_names, self.subtask_complete = None, [], False

    def select_action(self, obs):  # :82-104
        from .. import delegation_planner
        me = next(a for a in obs.sim_agents if a.name == self.name)
        self.location, self.holding, self.action = me.location, me.holding, me.action
        if obs.t == 0:
            self.setup_subtasks(env=obs)
        self.update_subtasks(obs)
        self.new_subtask, self.new_subtask_agent_names = self.delegator.select_subtask(agent_name=self.name)
        self.plan(obs)
        return self.action

    def __str__(self):  # :60-61
        return self.name[-1]

    def get_subtasks(self, world):  # :110-121 - the level's subtask list from the host recipe planner
        from .. import recipe_planner
        kinds = [nm for nm in ("Tomato", "Lettuce", "Onion", "Plate") for o in world.get_object_list()
                 if getattr(o, "mask", None) is not None and o.name == nm]
        return recipe_planner.level_subtasks(self.recipes, kinds, int(getattr(self.arglist, "max_num_subtasks", 14)))

    def setup_subtasks(self, env):  # :123-146
        from .. import delegation_planner
        self.incomplete_subtasks = (list(env.all_subtasks) if getattr(env, "all_subtasks", None) is not None
                                    else self.get_subtasks(world=env.world))
        self.delegator = delegation_planner.BayesianDelegator(
            agent_name=self.name, all_agent_names=env.get_agent_names(), model_type=self.model_type,
            planner=self.planner, none_action_prob=self.none_action_prob)

    def def_subtask_completion(self, env):  # :286-368
        from ..navigation_planner import goal_count
        from ..recipe_planner import subtask_masks
        from .core import Object
        goal_obj = Object((None, None), subtask_masks(self.new_subtask)[3])
        delivery = [gs.location for gs in env.world.objects.get("Delivery", [])]
        subtask = self.new_subtask
        base = goal_count(env.world, subtask, goal_obj, delivery)
        self.start_obj, self.goal_obj, self.cur_obj_count = None, goal_obj, base
        self.is_subtask_complete = lambda w: goal_count(w, subtask, goal_obj, delivery) > base

    def update_subtasks(self, env):  # :176-203
        if ((self.subtask is not None and self.subtask not in self.incomplete_subtasks)
                or self.delegator.should_reset_priors(obs=env, incomplete_subtasks=self.incomplete_subtasks)):
            self.reset_subtasks()
            self.delegator.set_priors(obs=env, incomplete_subtasks=self.incomplete_subtasks, priors_type=self.priors)
        elif self.subtask is None:
            self.delegator.set_priors(obs=env, incomplete_subtasks=self.incomplete_subtasks, priors_type=self.priors)
        else:
            self.delegator.bayes_update(obs_tm1=env.obs_tm1, actions_tm1=env.agent_actions, beta=self.beta)

    def refresh_subtasks(self, world):  # :151-171
        self.subtask_complete = False
        if self.subtask is None or len(self.subtask_agent_names) == 0:
            return
        self.subtask_complete = self.is_subtask_complete(world)
        if self.subtask_complete and self.subtask in self.incomplete_subtasks:
            self.incomplete_subtasks.remove(self.subtask)

    def all_done(self):
        from ..recipe_planner import Deliver
        return not any(isinstance(t, Deliver) for t in self.incomplete_subtasks)

    def plan(self, env):  # :218-281
        import numpy as np
        from ..delegation_planner import _single_actions
        if self.new_subtask is not None:
            self.def_subtask_completion(env=env)
        if self.new_subtask is None or not self.new_subtask_agent_names:
            me = next(a for a in env.sim_agents if a.name == self.name)
            actions = _single_actions(env, me)
            probs = [self.none_action_prob if a == (0, 0) else (1.0 - self.none_action_prob) / (len(actions) - 1)
                     for a in actions]
            self.action = actions[np.random.choice(len(actions), p=probs)]
        else:
            if self.model_type == "greedy":
                others = {}
            else:
                backup = self.new_subtask if self.new_subtask is not None else self.subtask
                others = self.delegator.get_other_agent_planners(obs=env, backup_subtask=backup)
            action = self.planner.get_next_action(env=env, subtask=self.new_subtask,
                                                  subtask_agent_names=self.new_subtask_agent_names,
                                                  other_agent_planners=others)
            if action is None:
                self.action = (0, 0)
            elif self.planner.is_joint:
                names = list(self.new_subtask_agent_names)
                self.action = action[names.index(self.name)] if self.name in names else action[0]
            else:
                self.action = action
        self.subtask, self.subtask_agent_names = self.new_subtask, self.new_subtask_agent_names
        self.new_subtask, self.new_subtask_agent_names = None, []
