"""Read-only World view over one packed env state (mirrors utils/world.py of the reference)."""
from collections import defaultdict

from .core import CELL_CLASS, AgentCounter, Delivery, GridSquare, Object


class World:
    NAV_ACTIONS = [(0, 1), (0, -1), (-1, 0), (1, 0)]  # world.py:16

    def __init__(self, level, arglist=None):
        self.arglist = arglist
        self.width, self.height = level.width, level.height
        self.perimeter = 2 * (self.width + self.height)  # env:198
        self.objects = defaultdict(list)
        self.loc_to_gridsquare = {}
        for y in range(self.height):
            for x in range(self.width):
                gs = CELL_CLASS[level.cell_type[y * 8 + x]]((x, y))
                self.objects[gs.name].append(gs)
                self.loc_to_gridsquare[(x, y)] = gs

    def set_objects(self, objects):
        """objects: [(mask, x, y, holder)] from engine.decode_state."""
        for name in [k for k, v in self.objects.items() if v and isinstance(v[0], Object)]:
            del self.objects[name]
        for gs in self.loc_to_gridsquare.values():
            gs.holding = [] if isinstance(gs, Delivery) else None
        for mask, x, y, holder in objects:
            obj = Object((x, y), mask, is_held=bool(holder))
            self.objects[obj.name].append(obj)
            if not holder:
                gs = self.loc_to_gridsquare[(x, y)]
                if isinstance(gs, Delivery):
                    gs.holding.append(obj)
                else:
                    gs.holding = obj

    def replace_with_agent_counter(self, location):
        """level-0 planning view (e2e_brtdp.py:405-406)."""
        old = self.loc_to_gridsquare[location]
        self.objects[old.name].remove(old) if old in self.objects[old.name] else None
        ac = AgentCounter(location)
        self.objects[ac.name].append(ac)
        self.loc_to_gridsquare[location] = ac

    def get_repr(self):
        return self.get_dynamic_objects()

    def update_display(self):  # world.py:42-53
        self.rep = [[" " for _ in range(self.width)] for _ in range(self.height)]
        show = lambda o: o.rep_str() if hasattr(o, "rep_str") else str(o)
        for obj in self.get_object_list():
            x, y = obj.location
            self.rep[y][x] = show(obj)
        for obj in self.objects.get("Tomato", []):
            x, y = obj.location
            self.rep[y][x] = show(obj)
        return self.rep

    def print_objects(self):  # world.py:55-58
        for k, v in self.objects.items():
            print(k, [o.location for o in v])

    def get_object_list(self):
        out = []
        for v in self.objects.values():
            out += v
        return out

    def get_dynamic_objects(self):  # world.py:323-337
        objs = []
        for key in sorted(self.objects.keys()):
            if key not in ("Counter", "Floor", "Delivery", "Cutboard") and "Supply" not in key:
                objs.append(tuple(o.get_repr() for o in self.objects[key]))
        return tuple(objs)

    def is_occupied(self, location):
        return any(isinstance(o, Object) and o.location == location and not o.is_held
                   for o in self.get_object_list())

    def get_object_locs(self, obj, is_held):  # world.py:354-375
        if obj.name not in self.objects:
            return []
        if isinstance(obj, Object):
            return [o.location for o in self.objects[obj.name] if obj == o and o.is_held == is_held]
        return [o.location for o in self.objects[obj.name] if obj == o]

    def get_all_object_locs(self, obj):
        return list(set(self.get_object_locs(obj, True) + self.get_object_locs(obj, False)))

    def get_object_at(self, location, desired_obj, find_held_objects):
        objs = [o for o in self.get_object_list()
                if isinstance(o, Object) and o.location == location and o.is_held is find_held_objects
                and (desired_obj is None or o.name == desired_obj.name)]
        assert len(objs) == 1, "looking for %s, found %d at %s" % (desired_obj, len(objs), location)
        return objs[0]

    def get_gridsquare_at(self, location):
        return self.loc_to_gridsquare[location]

    def inbounds(self, location):
        x, y = location
        return min(max(x, 0), self.width - 1), min(max(y, 0), self.height - 1)

    def is_collidable(self, location):
        return self.loc_to_gridsquare[location].collidable
