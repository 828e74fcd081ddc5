"""The nine kitchens of the reference, composed from their two parameters.

The reference ships them as `utils/levels/<divider>-divider_<recipe>.txt` (7 map rows, blank
line, recipe class names, blank line, agent `x y` lines; format parsed by
`OvercookedEnvironment.load_level`, envs/overcooked_environment.py:130-198).  All nine share
one 7x7 room with a tomato, a lettuce and two plates on the right-hand counters, two
cutboards and a delivery square on the left wall; they differ only in the divider column
(x = 3) and in the recipe list.  `level_text` rebuilds the text of a named level; any other
file in the same format can be loaded with `load_level_file`.
"""
import os

DIVIDERS = ("open", "partial", "full")
RECIPES = {
    "tomato": ("SimpleTomato",),
    "tl": ("SimpleTomato", "SimpleLettuce"),
    "salad": ("Salad",),
}
LEVEL_NAMES = tuple("%s-divider_%s" % (d, r) for d in DIVIDERS for r in ("tomato", "tl", "salad"))
AGENT_STARTS = ((2, 1), (4, 1), (4, 4), (2, 4))

# left wall (x = 0) and right wall (x = 6) for the five interior rows y = 1..5
_LEFT = "//*--"
_RIGHT = "l---p"
# rows whose divider square (x = 3) is a counter
_DIVIDER_ROWS = {"open": (), "partial": (1, 2, 3, 4), "full": (1, 2, 3, 4, 5)}


def level_text(name):
    """Text of the reference level `name`, e.g. 'partial-divider_tl'."""
    try:
        divider, recipe = name.split("-divider_")
        walls = _DIVIDER_ROWS[divider]
        recipes = RECIPES[recipe]
    except (ValueError, KeyError):
        raise KeyError("unknown level %r (known: %s)" % (name, ", ".join(LEVEL_NAMES)))
    rows = ["-----t-"]
    for y in range(1, 6):
        mid = "  -  " if y in walls else "     "
        rows.append(_LEFT[y - 1] + mid + _RIGHT[y - 1])
    rows.append("-----p-")
    agents = ["%d %d" % xy for xy in AGENT_STARTS]
    return "\n".join(rows + [""] + list(recipes) + [""] + agents) + "\n"


def load_level_file(path):
    with open(path, "r") as f:
        return f.read()


def resolve_level(level):
    """`arglist.level` -> level text: a known name, a path, or `utils/levels/<name>.txt`
    relative to the cwd exactly as the reference opens it (env:146)."""
    if level in LEVEL_NAMES:
        return level_text(level)
    for cand in (level, os.path.join("utils", "levels", "%s.txt" % level)):
        if os.path.isfile(cand):
            return load_level_file(cand)
    raise FileNotFoundError("level %r is neither a built-in level nor a file" % (level,))
