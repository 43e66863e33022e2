import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree (build container only)")


def _has_gpu():
    """Ask the product library itself (te_device_count): most GPU tests and the product do not need torch, and a box
    with a CPU-only torch must not silently skip the only guard of the bit-exact arithmetic."""
    try:
        from traffic_env_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    has_gpu = None
    expr = (config.getoption("markexpr") or "").replace(" ", "")
    if "gpu" in expr and "notgpu" not in expr and not _has_gpu():
        # `-m gpu` was asked for explicitly: a green run of nothing but skips would be a lie
        raise pytest.UsageError("-m gpu requested but libtraffic_b200.so is missing or sees no CUDA device "
                                "(build with `python __graft_entry__.py`; there is no CPU fallback)")
    from oracle import ref_harness
    has_ref = ref_harness.available()
    for item in items:
        if "gpu" in item.keywords:
            if has_gpu is None:
                has_gpu = _has_gpu()
            if not has_gpu:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not present"))


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
