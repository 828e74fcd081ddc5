from .overcooked_environment import OvercookedEnvironment, BatchObs, EnvView, CollisionRepr  # noqa: F401
