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

    def __repr__(self):
        return "SimAgent(%s @%s holding %s)" % (self.name, self.location, self.get_holding())
