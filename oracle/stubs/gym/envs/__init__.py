import importlib
from . import registration  # noqa: F401
from .registration import register  # noqa: F401


def make(env_id, **kwargs):
    if ":" in env_id:
        mod, env_id = env_id.split(":")
        importlib.import_module(mod)
    entry = registration.REGISTRY[env_id]
    mod_name, cls_name = entry.split(":")
    return getattr(importlib.import_module(mod_name), cls_name)(**kwargs)
