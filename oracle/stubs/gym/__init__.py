"""Minimal stand-in for gym 0.17.3 (not installable here): just enough for the
reference's gym_graph package to register and instantiate its environments.
Test infrastructure only -- see oracle/README.md."""
import importlib

from . import error, spaces, utils  # noqa: F401

_REG = {}


class Env(object):
    pass


def make(env_id):
    mod, cls = _REG[env_id].split(":")
    return getattr(importlib.import_module(mod), cls)()
