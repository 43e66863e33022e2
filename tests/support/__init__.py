"""TEST SUPPORT - not product code.

The GPU box has neither the reference checkout nor an old-API `gym`, so the tests bring minimal stand-ins for the
two third-party / out-of-scope modules the drop-in `gym_traffic` package imports (`gym_compat/gym`: the `_step/_reset`
gym API the reference is written against; `args_compat/args.py`: the reference's flag module), restatements of the
reference launcher's host-side wrapper shims (`wrappers_ref.py`: Warmup / History / Localize / Squish / UnGSpace and
`make_env`, traffic_test.py:66-91, gym_traffic/wrappers/*.py) and of its TensorFlow-free baseline controllers
(`run_baselines.py`).  In the build container the reference's own files are used instead
(tests/test_reference_live.py).
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
GYM_COMPAT = os.path.join(HERE, "gym_compat")
ARGS_COMPAT = os.path.join(HERE, "args_compat")


def _importable(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install_dropin(reference_dir=None):
    """traffic_env_b200.install.install() with the stand-in `gym` / `args` modules filled in where missing."""
    if reference_dir and reference_dir not in sys.path:
        sys.path.append(reference_dir)
    if not _importable("gym") or not hasattr(sys.modules["gym"], "Env") or not hasattr(sys.modules["gym"].Env, "_step"):
        for k in [k for k in sys.modules if k == "gym" or k.startswith("gym.")]:
            del sys.modules[k]
        sys.path.insert(0, GYM_COMPAT)
    if not _importable("args"):
        sys.path.insert(0, ARGS_COMPAT)
    import traffic_env_b200.install as inst
    return inst.install(reference_dir=reference_dir)
