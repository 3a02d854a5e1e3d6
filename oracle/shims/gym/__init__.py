"""Minimal stand-in for gym 0.21 so the UNMODIFIED reference imports headless.

Test infrastructure only (oracle harness). Only the names the reference touches
at import/constructor time exist (envs/drone_v2.py:1,10,77,120,134-149;
envs/__init__.py:1). Nothing here computes anything.
"""
from . import spaces, logger  # noqa: F401
from .envs import registration  # noqa: F401

_REGISTRY = {}


class Env(object):
    pass


def make(env_id, **kwargs):
    import importlib
    entry = _REGISTRY[env_id]
    mod, cls = entry.split(":")
    return getattr(importlib.import_module(mod), cls)(**kwargs)
