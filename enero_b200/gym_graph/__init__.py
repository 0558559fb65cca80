"""Drop-in for the reference's gym-graph package (gym-graph/gym_graph/__init__.py:4-17):
``GraphEnv-v16`` (DRL middlepoint environment), ``GraphEnv-v15`` (Local-Search environment) and
``GraphEnv-v20`` (SAP baseline environment) backed by the device-resident episode engine.  ``make(id)`` mirrors ``gym.make``; if the real
``gym`` is importable the ids are registered there too."""
from .envs import Env15, Env16, Env20

_REGISTRY = {"GraphEnv-v15": Env15, "GraphEnv-v16": Env16, "GraphEnv-v20": Env20}


def make(env_id, **kwargs):
    if env_id not in _REGISTRY:
        raise KeyError("%s is not provided by enero_b200" % env_id)
    return _REGISTRY[env_id](**kwargs)


try:  # pragma: no cover - gym is not installed in the build image
    from gym.envs.registration import register as _register
    for _id, _cls in (("GraphEnv-v15", "Env15"), ("GraphEnv-v16", "Env16"), ("GraphEnv-v20", "Env20")):
        try:
            _register(id=_id, entry_point="enero_b200.gym_graph.envs:" + _cls)
        except Exception:
            pass
except Exception:
    pass
