"""Put the drop-in `gym_traffic` package (and, when missing, compat `gym` / `args`) on sys.path.

    import traffic_env_b200.install as inst; inst.install()
    import gym, gym_traffic            # gym_traffic.envs.TrafficEnv now steps on the B200

With the reference checkout also on sys.path (after ours), its wrappers/ and algorithms/ stay
importable as gym_traffic.wrappers / gym_traffic.algorithms: only the simulator modules
(gym_traffic/__init__.py, envs/, spaces/) are replaced.
"""
import importlib
import os
import sys

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
DROPIN = os.path.join(_SHIM, "dropin")
GYM_COMPAT = os.path.join(_SHIM, "gym_compat")
ARGS_COMPAT = os.path.join(_SHIM, "args_compat")


def _importable(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install(reference_dir=None):
    """reference_dir: optional path of a traffic-env checkout whose args.py / wrappers / algorithms to reuse."""
    if reference_dir and reference_dir not in sys.path:
        sys.path.append(reference_dir)
    if reference_dir:
        os.environ["TRAFFIC_ENV_REFERENCE"] = reference_dir
    if not _importable("gym") or not hasattr(sys.modules["gym"], "Env") or not hasattr(sys.modules["gym"].Env, "_step"):
        for k in [k for k in sys.modules if k == "gym" or k.startswith("gym.")]:
            del sys.modules[k]
        sys.path.insert(0, GYM_COMPAT)
    if not _importable("args"):
        sys.path.insert(0, ARGS_COMPAT)
    for k in [k for k in sys.modules if k == "gym_traffic" or k.startswith("gym_traffic.")]:
        del sys.modules[k]
    if DROPIN in sys.path:
        sys.path.remove(DROPIN)
    sys.path.insert(0, DROPIN)
    import numpy as np
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_  # numpy 2 removed the alias the reference's callers may still use
    return DROPIN
