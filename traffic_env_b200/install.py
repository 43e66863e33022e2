"""Put the drop-in `gym_traffic` package on sys.path, ahead of any other `gym_traffic`.

    import traffic_env_b200.install as inst; inst.install(reference_dir="/path/to/traffic-env")
    import gym, gym_traffic            # gym_traffic.envs.TrafficEnv now steps on the B200

Only the simulator modules are replaced (gym_traffic/__init__.py, envs/, spaces/).  With the reference
checkout given (or already on sys.path) its args.py, wrappers/ and algorithms/ are reused unchanged:
`gym_traffic.wrappers.*` / `gym_traffic.algorithms.*` resolve there.  `gym` (the old `_step/_reset` API the
reference is written against) and the reference's `args` module must be importable; this package does not
ship stand-ins for them (the test-suite has its own under tests/support/, because the GPU box has neither).
"""
import importlib
import os
import sys

DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim", "dropin")


def install(reference_dir=None):
    """reference_dir: optional path of a traffic-env checkout whose args.py / wrappers / algorithms to reuse."""
    if reference_dir and reference_dir not in sys.path:
        sys.path.append(reference_dir)
    if reference_dir:
        os.environ["TRAFFIC_ENV_REFERENCE"] = reference_dir
    for name, what in (("gym", "an old-API gym (gym.Env with _step/_reset)"), ("args", "the reference's args.py")):
        try:
            importlib.import_module(name)
        except Exception as ex:
            raise ImportError("traffic_env_b200.install: %s must be importable (%s: %s)" % (what, type(ex).__name__, ex))
    for k in [k for k in sys.modules if k == "gym_traffic" or k.startswith("gym_traffic.")]:
        del sys.modules[k]
    if DROPIN in sys.path:
        sys.path.remove(DROPIN)
    sys.path.insert(0, DROPIN)
    import numpy as np
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_  # numpy 2 removed the alias the reference's callers may still use
    return DROPIN
