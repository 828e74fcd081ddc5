"""Read-only host views of GridSquares and Objects, materialised from the packed state.

Mirrors the data model of the reference's utils/core.py (GridSquare family :28-120, Object
:130-219, Food/Plate :254-366, mergeable :222-241) closely enough for its callers
(RealAgent, planners, Bag) to read: names, full names, locations, held flags, contents and
the predicates.  All mutation happens in the CUDA kernels; these objects never write back.
"""
from collections import namedtuple

GridSquareRepr = namedtuple("GridSquareRepr", "name location holding")
ObjectRepr = namedtuple("ObjectRepr", "name location is_held")

M_TOMATO, M_LETTUCE, M_ONION, M_PLATE = 1, 2, 4, 8
KINDS = (("Lettuce", M_LETTUCE), ("Onion", M_ONION), ("Plate", M_PLATE), ("Tomato", M_TOMATO))  # alphabetical


class Rep:
    FLOOR, COUNTER, CUTBOARD, DELIVERY = " ", "-", "/", "*"
    TOMATO, LETTUCE, ONION, PLATE = "t", "l", "o", "p"


class GridSquare:
    collidable = True

    def __init__(self, name, location):
        self.name = name
        self.location = location
        self.holding = None
        self.dynamic = False

    def __eq__(self, o):
        return isinstance(o, GridSquare) and self.name == o.name

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return "%s%r" % (self.name, (self.location,))

    def __str__(self):  # core.py:37-38, without the terminal colour codes
        return self.rep


class Floor(GridSquare):
    collidable = False
    rep = Rep.FLOOR

    def __init__(self, location):
        GridSquare.__init__(self, "Floor", location)


class Counter(GridSquare):
    rep = Rep.COUNTER

    def __init__(self, location):
        GridSquare.__init__(self, "Counter", location)


class AgentCounter(Counter):
    def __init__(self, location):
        GridSquare.__init__(self, "Agent-Counter", location)

    def get_repr(self):
        return GridSquareRepr(name=self.name, location=self.location, holding=None)


class Cutboard(GridSquare):
    rep = Rep.CUTBOARD

    def __init__(self, location):
        GridSquare.__init__(self, "Cutboard", location)


class Delivery(GridSquare):
    rep = Rep.DELIVERY

    def __init__(self, location):
        GridSquare.__init__(self, "Delivery", location)
        self.holding = []


CELL_CLASS = {0: Floor, 1: Counter, 2: Cutboard, 3: Delivery}


class Food:
    def __init__(self, name, chopped):
        self.name = name
        self.state_index = 1 if chopped else 0
        self.full_name = ("Chopped" if chopped else "Fresh") + name

    def get_state(self):
        return "Chopped" if self.state_index else "Fresh"

    def needs_chopped(self):
        return self.state_index == 0

    def done(self):
        return self.state_index == 1

    def __eq__(self, other):
        return isinstance(other, Food) and self.get_state() == other.get_state()

    def __hash__(self):
        return hash(self.full_name)


class Plate:
    name = full_name = "Plate"

    def needs_chopped(self):
        return False

    def __eq__(self, other):
        return isinstance(other, Plate)

    def __hash__(self):
        return hash("Plate")


def mask_names(mask):
    """(name, full_name) of an object with this content mask (Object.update_names :161-171)."""
    names, full = [], []
    for kind, bit in KINDS:
        if mask & bit:
            names.append(kind)
            if kind == "Plate":
                full.append("Plate")
            else:
                full.append(("Chopped" if mask & (bit << 4) else "Fresh") + kind)
    return "-".join(names), "-".join(full)


def name_to_mask(full_name):
    m = 0
    if full_name in (None, "None", ""):
        return 0
    for part in full_name.split("-"):
        if part == "Plate":
            m |= M_PLATE
        else:
            chopped = part.startswith("Chopped")
            kind = part[7:] if chopped else part[5:] if part.startswith("Fresh") else part
            bit = dict(KINDS)[kind]
            m |= bit | ((bit << 4) if chopped else 0)
    return m


class Object:
    collidable = False

    def __init__(self, location, mask, is_held=False):
        self.location = location
        self.mask = mask
        self.is_held = is_held
        self.name, self.full_name = mask_names(mask)
        self.contents = [Plate() if k == "Plate" else Food(k, bool(mask & (b << 4)))
                         for k, b in KINDS if mask & b]

    def __eq__(self, other):  # :143-148
        return (isinstance(other, Object) and self.name == other.name
                and len(self.contents) == len(other.contents) and self.full_name == other.full_name)

    def __hash__(self):
        return hash(self.full_name)

    def __str__(self):
        return self.full_name

    __repr__ = __str__

    def rep_str(self):
        """the reference's Object.__str__ (core.py:139-141): content letters joined by '-', sorted by name"""
        return "-".join(nm[0].lower() for nm, bit in KINDS if self.mask & bit)

    def get_repr(self):
        return ObjectRepr(name=self.full_name, location=self.location, is_held=self.is_held)

    def contains(self, c_name):
        return c_name in [c.name for c in self.contents]

    def needs_chopped(self):
        return self.mask in (1, 2, 4)

    def is_chopped(self):
        return not (self.mask & M_PLATE) and foods_done(self.mask)

    def is_merged(self):
        return len(self.contents) > 1

    def is_deliverable(self):
        return self.is_merged() and foods_done(self.mask)


def foods_done(mask):
    return ((mask & 7) & ~(mask >> 4)) == 0


def mergeable(obj1, obj2):
    m1, m2 = obj1.mask, obj2.mask
    return not (m1 & m2 & M_PLATE) and foods_done(m1 | m2)
