"""gym-cooking_b200: B200-native batched drop-in for the three hot paths of deletfsi/gym-cooking
(env step, subtask value/Q, Bayesian-Delegation posterior).  See DESIGN.md."""
from . import levels  # noqa: F401
from . import _lib  # noqa: F401
from .engine import KitchenBatch, ACTIONS, ACTION_INDEX, decode_state  # noqa: F401
from .envs import OvercookedEnvironment  # noqa: F401,E402

ENV_ID = "overcookedEnv-v0"  # gym id the reference registers (gym_cooking/__init__.py:3-6)


def make(env_id="gym_cooking:overcookedEnv-v0", arglist=None, **kwargs):
    """Stand-in for `gym.envs.make("gym_cooking:overcookedEnv-v0", arglist=arglist)` (main.py:88);
    if `gym` is installed the same id is also registered with it on import."""
    if env_id.split(":")[-1] != ENV_ID:
        raise ValueError("unknown env id %r" % (env_id,))
    return OvercookedEnvironment(arglist, **kwargs)


try:  # pragma: no cover - gym is not in this image
    from gym.envs.registration import register as _register
    _register(id=ENV_ID, entry_point="gym_cooking_b200.envs:OvercookedEnvironment")
except Exception:
    pass
from .planning import bd_likelihood_rows, bd_posterior, lower_bound, subtask_q, subtask_q_unique  # noqa: F401,E402
