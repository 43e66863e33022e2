"""CPU: the oracle (oracle/traffic_oracle.c) against the golden vectors that
oracle/gen_golden.py produced by running the unmodified reference.  This is the
pin that lets the GPU parity tests trust the oracle on the GPU box, where the
reference itself cannot run."""
import glob
import os

import numpy as np
import pytest

from oracle.oracle import OracleEnv
from tests.golden_util import tick_digest, unpack_schedule

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RAW = sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if str(np.load(p)["kind"]) == "raw")
WRAPPED = sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if str(np.load(p)["kind"]) == "wrapped")


def make_oracle(g):
    o = OracleEnv(int(g["m"]), int(g["n"]), float(g["length"]), float(g["rate"]),
                  learn_switch=bool(g["learn_switch"]) if "learn_switch" in g else False,
                  validate=bool(g["validate"]) if "validate" in g else False)
    if "entry_spec" in g:
        o.generate_entrypoints(int(g["entry_spec"]))
    return o


def test_fixture_inventory():
    names = {os.path.basename(p) for p in RAW + WRAPPED}
    assert {"kat_fixed_3x3.npz", "overflow_3x3.npz", "grid10_len500.npz", "wrapped_3x3_cfg2.npz"} <= names


@pytest.mark.parametrize("path", RAW, ids=[os.path.basename(p)[:-4] for p in RAW])
def test_raw_ticks_bit_exact(path):
    g = np.load(path)
    o = make_oracle(g)
    assert (o.entrypoints == g["entrypoints"]).all()
    o.reset(g["init_phase"])
    sched = unpack_schedule(g["sched_off"], g["sched_roads"])
    T = int(g["ticks"])
    for t in range(T):
        done = o.step(g["actions"][t], sched[t])
        xs, vs = o.live_state()
        d = tick_digest(o.leading, o.lastcar, o.obs, o.waiting, o.passed_dst, o.rewards, done, xs, vs)
        assert d == g["digests"][t], "first divergent tick %d" % t
        assert done == bool(g["dones"][t])
        assert (o.rewards == g["rewards"][t]).all()
        key = "ck%d_x" % (t + 1)
        if key in g:
            assert xs.tobytes() == g[key].tobytes()
            assert vs.tobytes() == g["ck%d_v" % (t + 1)].tobytes()
            assert (o.leading == g["ck%d_leading" % (t + 1)]).all()
            assert (o.obs == g["ck%d_obs" % (t + 1)]).all()
    assert o.generated_cars == int(g["generated"][-1])
    if "trip_times" in g:
        assert o.trip_times().tobytes() == g["trip_times"].tobytes()


def test_kat_hashes_from_survey():
    """first 16 hex of sha256 over raw bytes, KAT-B (200) / KAT-A (1200)."""
    import hashlib
    g = np.load(os.path.join(GOLDEN, "kat_fixed_3x3.npz"))
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
    assert list(g["init_phase"]) == [0, 1, 1, 0, 1, 1, 1, 1, 1]
    assert h(g["ck200_leading"]) == "db62d6f5912add19" and h(g["ck200_x"]) == "bbac51daa0751502"
    assert h(g["ck200_v"]) == "102f5680c79c3077" and h(g["ck200_obs"]) == "dfa3cf3af708728b"
    assert h(g["ck1200_leading"]) == "e94dceb59dc64b2b" and h(g["ck1200_lastcar"]) == "430992a7a6402fda"
    assert h(g["ck1200_obs"]) == "daf5c7f8d3da3a71" and h(g["ck1200_x"]) == "13cb79d775c93a9c"
    assert h(g["ck1200_v"]) == "4a507ec168da5bea"
    assert int(g["generated"][199]) == 134 and int(g["generated"][1199]) == 859


def oracle_actor_step(o, action, sched, t0, K, use_remi=True):
    """Remi(Repeater(K)) restated on the oracle (traffic_test.py:37-64)."""
    r, I = o.train_roads, o.intersections
    obs = np.zeros(2 * r + I, dtype=np.float32)
    total = np.zeros(I, dtype=np.float32)
    done = False
    n = 0
    for k in range(K):
        done = o.step(action, sched[t0 + k] if t0 + k < len(sched) else ())
        n += 1
        obs[:r] += o.obs[:r]
        obs[r:2 * r] = o.obs[r:2 * r]
        mult = 2 * o.obs[-2 * I:-I] - 1
        obs[-I:] = o.obs[-I:] / 100 * mult
        total += o.rewards
        if done:
            break
    rew = o.remi_reward().copy() if use_remi else total
    return obs, rew, done, n


@pytest.mark.parametrize("path", WRAPPED, ids=[os.path.basename(p)[:-4] for p in WRAPPED])
def test_wrapped_actor_steps_bit_exact(path):
    g = np.load(path)
    K, S, E = int(g["K"]), int(g["actor_steps"]), int(g["n_envs"])
    for e in range(E):
        o = make_oracle(g)
        o.reset(g["init_phase"][e])
        lo, hi = g["sched_roads_off"][e], g["sched_roads_off"][e + 1]
        sched = unpack_schedule(g["sched_off"][e], g["sched_roads"][lo:hi])
        # The reference generator is consumed once per tick actually run (Repeater
        # breaks on done), so the schedule cursor advances by ticks run, not by K.
        t = 0
        for s in range(S):
            obs, rew, done, n = oracle_actor_step(o, g["actions"][s, e], sched, t, K)
            t += n
            assert obs.tobytes() == g["obs"][e, s].tobytes(), (e, s)
            assert rew.tobytes() == g["reward"][e, s].tobytes(), (e, s)
            assert done == bool(g["done"][e, s])
        xs, vs = o.live_state()
        assert xs.tobytes() == g["fin%d_x" % e].tobytes() and vs.tobytes() == g["fin%d_v" % e].tobytes()
        assert (o.leading == g["fin%d_leading" % e]).all() and (o.lastcar == g["fin%d_lastcar" % e]).all()
        assert (o.obs == g["fin%d_obs" % e]).all()
