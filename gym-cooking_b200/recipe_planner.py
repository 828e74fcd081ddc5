"""Host-side recipe planner: which subtasks a level's recipes need.

Out of scope as a kernel (runs once per level) but subtask IDENTITY is
part of the navigation/delegation API, so the `Get/Chop/Merge/Deliver` value objects and the
"union of actions over all shortest STRIPS plans" rule of the reference are kept
(recipe_planner/utils.py:63-162, recipe.py:5-228, stripsworld.py:25-93).  The search itself is
a plain layered BFS over predicate multisets with a backward sweep that marks every action on
some shortest plan.  Subtasks are returned in a deterministic order (the reference iterates a
Python set, whose order depends on PYTHONHASHSEED).
"""
from itertools import combinations

from .utils.core import name_to_mask

FOODS = ("Tomato", "Lettuce", "Onion")
RECIPE_INGREDIENTS = {
    "SimpleTomato": ("Tomato",),
    "SimpleLettuce": ("Lettuce",),
    "Salad": ("Tomato", "Lettuce"),
    "OnionSalad": ("Tomato", "Lettuce", "Onion"),
}
ST_NONE, ST_CHOP, ST_MERGE, ST_DELIVER = 0, 1, 2, 3


class Action:
    """A recipe subtask, identified by (name, args) exactly like the reference's Action."""
    name = "Action"

    def __init__(self, *args, pre=None, post=None):
        self.args = tuple(args)
        self.pre = list(pre) if pre is not None else self.default_pre()
        self.post_add = list(post) if post is not None else self.default_post()
        self.is_joint = False

    def __eq__(self, other):
        return other is not None and getattr(other, "name", None) == self.name and getattr(other, "args", None) == self.args

    def __hash__(self):
        return hash((self.name, self.args))

    def __str__(self):
        return "%s(%s)" % (self.name, ", ".join(self.args))

    __repr__ = __str__


def _join(*names):
    parts = []
    for n in names:
        parts += n.split("-")
    return "-".join(sorted(parts))


class Get(Action):
    name = "Get"

    def default_pre(self):
        return [("None", None)]

    def default_post(self):
        return [("Fresh", self.args[0]), ("None", None)]


class Chop(Action):
    name = "Chop"

    def default_pre(self):
        return [("Fresh", self.args[0])]

    def default_post(self):
        return [("Chopped", self.args[0])]


class Merge(Action):
    name = "Merge"

    def default_pre(self):
        return [("Chopped", self.args[0]), ("Merged", self.args[1])]

    def default_post(self):
        return [("Merged", _join(*self.args))]


class Deliver(Action):
    name = "Deliver"

    def default_pre(self):
        return [("Merged", self.args[0])]

    def default_post(self):
        return [("Delivered", self.args[0])]


def recipe_actions(recipe_name):
    """Action set and goal predicate of a recipe class (recipe.py:5-228)."""
    foods = sorted(RECIPE_INGREDIENTS[recipe_name])
    acts = {Get("Plate")}
    for f in foods:
        acts.add(Get(f))
        acts.add(Chop(f))
        acts.add(Merge(f, "Plate", pre=[("Chopped", f), ("Fresh", "Plate")]))
    plated = _join(*foods, "Plate")
    acts.add(Deliver(plated))
    for size in range(2, len(foods) + 1):
        for combo in combinations(foods, size):
            whole = _join(*combo)
            acts.add(Merge(whole, "Plate", pre=[("Merged", whole), ("Fresh", "Plate")]))
            for item in combo:
                rest = [c for c in combo if c != item]
                rest_name, item_plate, rest_plate = _join(*rest), _join(item, "Plate"), _join(*rest, "Plate")
                if len(rest) == 1:
                    acts.add(Merge(item, rest_name, pre=[("Chopped", item), ("Chopped", rest_name)]))
                    acts.add(Merge(rest_name, item_plate))
                    acts.add(Merge(item, rest_plate))
                else:
                    acts.add(Merge(item, rest_name))
                    acts.add(Merge(item_plate, rest_name, pre=[("Merged", item_plate), ("Merged", rest_name)]))
                    acts.add(Merge(item, rest_plate))
    return acts, ("Delivered", plated)


def _apply(state, action):
    """state: sorted tuple of predicates (a multiset).  None if a precondition is missing."""
    s = list(state)
    for p in action.pre:
        if p not in s:
            return None
        s.remove(p)
    s += action.post_add
    return tuple(sorted(s, key=repr))


def shortest_plan_actions(initial, actions, goal, max_depth=14):
    """All actions that lie on some shortest plan from `initial` to the FIRST goal state found
    (stripsworld.generate_graph :25-62 returns at the first hit; get_subtasks :68-93 unions the
    edge labels of nx.all_shortest_paths to it)."""
    actions = sorted(actions, key=str)
    depth = {initial: 0}
    parents = {initial: []}
    frontier = [initial]
    goal_state = None
    for d in range(max_depth):
        nxt = []
        for st in frontier:
            for a in actions:
                ns = _apply(st, a)
                if ns is None:
                    continue
                if ns not in depth:
                    depth[ns] = d + 1
                    parents[ns] = []
                    nxt.append(ns)
                # one edge per (state, next state): the reference's nx.DiGraph keeps a single
                # action per edge (stripsworld.py:47, the later add_edge overwrites `obj`), so of
                # two actions with identical effects - Merge(Lettuce, Tomato) / Merge(Tomato,
                # Lettuce) - only one survives, WHICH one depends on its set order; we keep the
                # lexicographically first
                if depth[ns] == d + 1 and all(ps != st for ps, _ in parents[ns]):
                    parents[ns].append((st, a))
                if goal in ns and goal_state is None:
                    goal_state = ns
        if goal_state is not None:
            break
        frontier = nxt
    if goal_state is None:
        raise ValueError("goal state could not be found, try increasing --max-num-subtasks")
    used, seen, stack = set(), {goal_state}, [goal_state]
    while stack:
        st = stack.pop()
        for (ps, a) in parents[st]:
            used.add(a)
            if ps not in seen:
                seen.add(ps)
                stack.append(ps)
    return used


_KIND_ORDER = {"Get": 0, "Chop": 1, "Merge": 2, "Deliver": 3}


def level_subtasks(recipe_names, object_kinds, max_num_subtasks=14):
    """env.run_recipes (env:396-473): per recipe, the union over all shortest plans, flattened.
    `object_kinds`: kinds of the objects lying in the level at reset (one Fresh predicate each,
    stripsworld.py:17-23)."""
    initial = tuple(sorted([("None", None)] + [("Fresh", k) for k in object_kinds], key=repr))
    out = []
    for r in recipe_names:
        acts, goal = recipe_actions(r)
        used = shortest_plan_actions(initial, acts, goal, max_num_subtasks)
        out += sorted(used, key=lambda a: (_KIND_ORDER[a.name], a.args))
    return out


def _object_mask(name):
    """mask of 'Lettuce-Plate-Tomato' style names with every food in its last (chopped) state."""
    m = 0
    for part in name.split("-"):
        m |= name_to_mask("Plate" if part == "Plate" else "Chopped" + part)
    return m


def subtask_masks(subtask):
    """(kind, a, b, goal) in content-mask form - nav_utils.get_subtask_obj :181-246."""
    if subtask is None:
        return (ST_NONE, 0, 0, 0)
    if isinstance(subtask, Chop):
        fresh = name_to_mask("Fresh" + subtask.args[0])
        return (ST_CHOP, fresh, 0, fresh | (fresh << 4))
    if isinstance(subtask, Merge):
        a, b = _object_mask(subtask.args[0]), _object_mask(subtask.args[1])
        return (ST_MERGE, a, b, a | b)
    if isinstance(subtask, Deliver):
        m = _object_mask(subtask.args[0])
        return (ST_DELIVER, m, 0, m)
    raise NotImplementedError("%s was not recognized" % (subtask,))
