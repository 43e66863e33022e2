"""CPU: the host-side restatement of the reference's arrival generators (traffic_env_b200/host_arrivals.py)
reproduces the schedules that oracle/gen_golden.py recorded from the reference's own generators
(traffic_env.py:160-176, 274-283 driven by RandomState(seed))."""
import os

import numpy as np
import pytest

from tests.golden_util import unpack_schedule
from traffic_env_b200.host_arrivals import ArrivalStream

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cars_per_sec(local, m, spec):
    return local * m * bin((~spec) & 0b1111).count("1")


@pytest.mark.parametrize("name,seed,local,poisson", [
    ("kat_fixed_3x3", 0, 0.12, True), ("overflow_3x3", 3, 0.9, True), ("overflow_2x2_stuck", 4, 1.2, True),
    ("validate_3x3", 5, 0.12, True), ("learnswitch_2x3", 6, 0.3, True), ("entry_one_3x3", 7, 0.5, True),
    ("regular_3x2", 8, 0.25, False), ("grid10_len500", 9, 0.35, True)])
def test_stream_reproduces_recorded_schedule(name, seed, local, poisson):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    spec = int(g["entry_spec"])
    want = unpack_schedule(g["sched_off"], g["sched_roads"])
    s = ArrivalStream(seed, g["entrypoints"], cars_per_sec(local, int(g["m"]), spec), float(g["rate"]), poisson=poisson)
    got = s.window(len(want))
    assert [list(map(int, w)) for w in want] == got


def test_stream_reproduces_config2_env_schedules():
    g = np.load(os.path.join(GOLDEN, "wrapped_3x3_cfg2.npz"))
    from traffic_env_b200.vec_env import VecTrafficEnv  # noqa: F401  (import only: no device needed)
    entry = np.array([0, 3, 6, 11, 14, 17, 18, 19, 20, 33, 34, 35])
    for e in range(int(g["n_envs"])):
        lo, hi = g["sched_roads_off"][e], g["sched_roads_off"][e + 1]
        want = unpack_schedule(g["sched_off"][e], g["sched_roads"][lo:hi])
        got = ArrivalStream(e, entry, cars_per_sec(0.12, 3, 0), 0.5).window(len(want))
        assert [list(map(int, w)) for w in want] == got
