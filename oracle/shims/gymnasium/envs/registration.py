"""register()/make() with entry-point strings (test infrastructure only)."""
import importlib

registry = {}


def register(id, entry_point=None, max_episode_steps=None, kwargs=None, **_):
    registry[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps, kwargs=dict(kwargs or {}))


def make(id, **kwargs):
    spec = registry[id]
    ep = spec["entry_point"]
    if isinstance(ep, str):
        mod, name = ep.split(":")
        ep = getattr(importlib.import_module(mod), name)
    kw = dict(spec["kwargs"])
    kw.update(kwargs)
    return ep(**kw)
