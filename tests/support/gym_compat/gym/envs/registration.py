"""Name -> entry point registry (the subset of gym.envs.registration used by
reference gym_traffic/__init__.py:20-23 and traffic_test.py:79)."""
import importlib


class EnvSpec(object):
    def __init__(self, id, entry_point=None, kwargs=None, **_ignored):
        self.id = id
        self.entry_point = entry_point
        self.kwargs = dict(kwargs or {})

    def make(self):
        if callable(self.entry_point):
            cls = self.entry_point
        else:
            mod_name, attr = self.entry_point.split(":")
            cls = getattr(importlib.import_module(mod_name), attr)
        env = cls(**self.kwargs)
        env.spec = self
        return env


class EnvRegistry(object):
    def __init__(self):
        self.env_specs = {}

    def register(self, id, **kwargs):
        # Re-registration replaces (the drop-in shim and the reference may both
        # register 'traffic-v0' in one interpreter during parity runs).
        self.env_specs[id] = EnvSpec(id, **kwargs)

    def spec(self, id):
        try:
            return self.env_specs[id]
        except KeyError:
            raise KeyError("No registered env with id: %s" % id)

    def make(self, id):
        return self.spec(id).make()


registry = EnvRegistry()


def register(id, **kwargs):
    return registry.register(id, **kwargs)


def make(id):
    return registry.make(id)


def spec(id):
    return registry.spec(id)
