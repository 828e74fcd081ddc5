"""gym-cooking_b200: B200-native batched drop-in for the three hot paths of deletfsi/gym-cooking
(env step, subtask value/Q, Bayesian-Delegation posterior).  See DESIGN.md."""
from . import levels  # noqa: F401
from . import _lib  # noqa: F401
from .engine import KitchenBatch, ACTIONS, ACTION_INDEX, decode_state  # noqa: F401
