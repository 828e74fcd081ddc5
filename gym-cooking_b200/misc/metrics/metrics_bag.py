"""`Bag`: per-episode record with the reference's pickle keys (misc/metrics/metrics_bag.py:5-72), so
`make_graphs.py` / `make_heatmap.py` read episodes produced here unchanged.

Keys: level, num_agents, profiling, num_completed_subtasks, agent-<i> (model types), states,
actions, subtasks, subtask_agents, bayes, holding, incomplete_subtasks, all_subtasks,
num_total_subtasks, collisions, termination, was_successful, num_completed_subtasks_end.
Differences: the directory is an argument (the reference hard-codes a Windows path, :9) and nothing
is printed on save."""
import copy
import os
import pickle


# per-agent series: key -> how one RealAgent contributes to it at a step (metrics_bag.py:44-51)
AGENT_SERIES = (
    ("states", lambda a: copy.copy(a.location)),
    ("holding", lambda a: a.get_holding()),
    ("actions", lambda a: a.action),
    ("subtasks", lambda a: a.subtask),
    ("subtask_agents", lambda a: a.subtask_agent_names),
    ("incomplete_subtasks", lambda a: a.incomplete_subtasks),
)


class Bag:
    def __init__(self, arglist, filename, directory="misc/metrics/pickles/"):
        self.arglist, self.filename, self.directory = arglist, filename, directory
        self.data = {}
        self.set_general()

    def set_general(self):  # :13-33
        names = ["agent-%d" % (i + 1) for i in range(self.arglist.num_agents)]
        self.data.update(level=self.arglist.level, num_agents=self.arglist.num_agents, num_completed_subtasks=[],
                         profiling={k: [] for k in ("Delegation", "Navigation", "Total")})
        for i in range(1, 5):  # model types of the ablation runs
            model = getattr(self.arglist, "model%d" % i, None)
            if model is not None:
                self.data["agent-%d" % i] = model
        for key, _ in AGENT_SERIES:
            self.data[key] = {nm: [] for nm in names}
        self.data["bayes"] = {nm: {} for nm in names}

    def set_recipe(self, recipe_subtasks):  # :36-38
        self.data.update(all_subtasks=recipe_subtasks, num_total_subtasks=len(recipe_subtasks))

    def set_collisions(self, collisions):  # :40-41
        self.data["collisions"] = collisions

    def add_status(self, cur_time, real_agents):  # :44-61
        still_open = set(self.data["all_subtasks"])
        for agent in real_agents:
            for key, value_of in AGENT_SERIES:
                self.data[key][agent.name].append(value_of(agent))
            beliefs = self.data["bayes"][agent.name].setdefault(cur_time, []) if agent.delegator.probs.get_list() else None
            for alloc, prob in agent.delegator.probs.get_list():
                beliefs.append((alloc, prob))
            still_open &= set(agent.incomplete_subtasks)  # open = what EVERY agent still believes open
        self.data["num_completed_subtasks"].append(self.data["num_total_subtasks"] - len(still_open))

    def set_termination(self, termination_info, successful, save=True):  # :63-72
        series = self.data["num_completed_subtasks"]
        self.data.update(termination=termination_info, was_successful=successful,
                         num_completed_subtasks_end=series[-1] if series else 0)
        return self.save() if save else None

    def save(self):
        os.makedirs(self.directory, exist_ok=True)
        path = os.path.join(self.directory, self.filename + ".pkl")
        with open(path, "wb") as f:
            pickle.dump(self.data, f)
        return path
