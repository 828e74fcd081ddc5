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


class Bag:
    def __init__(self, arglist, filename, directory="misc/metrics/pickles/"):
        self.data = {}
        self.arglist = arglist
        self.directory = directory
        self.filename = filename
        self.set_general()

    def set_general(self):  # :13-33
        n = self.arglist.num_agents
        self.data["level"] = self.arglist.level
        self.data["num_agents"] = n
        self.data["profiling"] = {info: [] for info in ["Delegation", "Navigation", "Total"]}
        self.data["num_completed_subtasks"] = []
        for i in range(1, 5):
            model = getattr(self.arglist, "model%d" % i, None)
            if model is not None:
                self.data["agent-%d" % i] = model
        for info in ["states", "actions", "subtasks", "subtask_agents", "bayes", "holding", "incomplete_subtasks"]:
            self.data[info] = {"agent-%d" % (i + 1): ({} if info == "bayes" else []) for i in range(n)}

    def set_recipe(self, recipe_subtasks):  # :36-38
        self.data["all_subtasks"] = recipe_subtasks
        self.data["num_total_subtasks"] = len(recipe_subtasks)

    def set_collisions(self, collisions):  # :40-41
        self.data["collisions"] = collisions

    def add_status(self, cur_time, real_agents):  # :44-61
        for a in real_agents:
            self.data["states"][a.name].append(copy.copy(a.location))
            self.data["holding"][a.name].append(a.get_holding())
            self.data["actions"][a.name].append(a.action)
            self.data["subtasks"][a.name].append(a.subtask)
            self.data["subtask_agents"][a.name].append(a.subtask_agent_names)
            self.data["incomplete_subtasks"][a.name].append(a.incomplete_subtasks)
            for task_combo, p in a.delegator.probs.get_list():
                self.data["bayes"][a.name].setdefault(cur_time, []).append((task_combo, p))
        incomplete = set(self.data["all_subtasks"])
        for a in real_agents:
            incomplete &= set(a.incomplete_subtasks)
        self.data["num_completed_subtasks"].append(self.data["num_total_subtasks"] - len(incomplete))

    def set_termination(self, termination_info, successful, save=True):  # :63-72
        self.data["termination"] = termination_info
        self.data["was_successful"] = successful
        done = self.data["num_completed_subtasks"]
        self.data["num_completed_subtasks_end"] = done[-1] if done else 0
        if save:
            return self.save()

    def save(self):
        os.makedirs(self.directory, exist_ok=True)
        path = os.path.join(self.directory, self.filename + ".pkl")
        with open(path, "wb") as f:
            pickle.dump(self.data, f)
        return path
