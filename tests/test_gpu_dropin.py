"""GPU: the drop-in gym_traffic package (TrafficEnv facade + wrappers) behaves like the reference's.

The reference tree is not on the GPU box, so the agents' loops (algorithms/fixed.py:9-23,
greedy.py:7-19, random.py:6-16) are restated here as their env-facing contract."""
import os
import sys

import numpy as np
import pytest

from tests.golden_util import live_walk, tick_digest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dropin():
    import traffic_env_b200.install as inst
    inst.install()
    import gym
    import gym_traffic  # noqa: F401
    from args import FLAGS
    from gym_traffic.envs.roadgraph import GridRoad
    FLAGS.local_cars_per_sec, FLAGS.rate, FLAGS.poisson, FLAGS.entry, FLAGS.learn_switch = 0.12, 0.5, True, "all", False
    return gym, GridRoad, FLAGS


def new_env(gym, GridRoad, m=3, n=3, length=250, seed=0):
    env = gym.make("traffic-v0")
    env.set_graph(GridRoad(m, n, length))
    env.seed_generator(seed)
    env.reset_entrypoints()
    return env


def test_kat_through_the_facade(dropin):
    """SURVEY.md 8c recipe, verbatim, on the drop-in: np.random.seed(0), seed_generator(0), fixed policy.
    The facade feeds the device the reference's own MT19937 arrival stream, so every tick digest of the
    golden run (ints and float bits) must match."""
    gym, GridRoad, FLAGS = dropin
    g = np.load(os.path.join(GOLDEN, "kat_fixed_3x3.npz"))
    np.random.seed(0)
    u = new_env(gym, GridRoad)
    u.reset()
    assert list(u.current_phase) == [0, 1, 1, 0, 1, 1, 1, 1, 1]
    for t in range(1200):
        a = np.ones(9, np.int32) if ((t // 10) % 6) >= 3 else np.zeros(9, np.int32)
        obs, rew, done, info = u.step(a)
        assert obs is u.obs and rew is u.rewards and info is None  # same buffers every call (traffic_env.py:248)
        if t % 50 == 49 or t < 20:
            u.sync_counters()
            st = u._snapshot()
            xs, vs = live_walk(st["leading"][0], st["lastcar"][0], st["x"][0], st["v"][0])
            d = tick_digest(st["leading"][0], st["lastcar"][0], obs, u.waiting, u.passed_dst, rew, done, xs, vs)
            assert d == g["digests"][t], "tick %d" % t
    assert u.generated_cars == 859 and float(u.steps) == 1200.0
    assert u.cars_on_roads().shape == (3, 3, 4)
    assert u.state.shape == (48, 10, 20)


def test_fused_wrappers_match_tick_loop(dropin):
    """Remi(Repeater(10)) with the fused single-launch Repeater == the same wrappers forced onto the
    reference's tick-by-tick loop (rendering flag disables fusion), action types as the agents send them."""
    gym, GridRoad, FLAGS = dropin
    from traffic_env_b200.wrappers import Remi, Repeater

    def build(fused):
        np.random.seed(3)
        base = new_env(gym, GridRoad, seed=11)
        if not fused:
            base.rendering = True           # per-tick rendering requested: Repeater must not fuse ...
            base.render = lambda *a, **k: None  # ... and the (absent) viewer is stubbed out
        return Remi(Repeater(10)(base)), base

    fa, fb = build(True)
    ta, tb = build(False)
    np.random.seed(21)
    o1 = fa.reset()
    np.random.seed(21)   # GSpace.sample() draws phases and the first action from the global RNG
    o2 = ta.reset()
    assert o1.dtype == np.float32 and o1.shape == (81,) and o1.tobytes() == o2.tobytes()
    rng = np.random.RandomState(0)
    sent = [lambda: rng.rand(9) < 0.5,                       # bool (a3c)
            lambda: rng.randint(2, size=9).astype(np.int32),  # int32 (dqn)
            lambda: rng.randint(2, size=9).astype(np.float64),  # float64 (fixed / const)
            lambda: rng.randint(2, size=9).astype(np.int8),   # int8 (cem)
            lambda: tuple(int(v) for v in rng.randint(2, size=9))]  # tuple (UnGSpaceWrapper)
    for s in range(40):
        a = sent[s % len(sent)]()
        x1, r1, d1, _ = fa.step(a)
        x2, r2, d2, _ = ta.step(a)
        assert x1.tobytes() == np.asarray(x2, np.float32).tobytes(), s
        assert np.asarray(r1).tobytes() == np.asarray(r2).tobytes() and d1 == d2, s
    assert fa.reward_size == 9 and fa.observation_space.shape == [81] and fa.action_space.size == 9


def test_agent_loop_contracts(dropin):
    gym, GridRoad, FLAGS = dropin
    from traffic_env_b200.wrappers import make_env
    FLAGS.light_iterations = 10
    env = make_env(seed=5)
    # fixed.py: float64 0./1. actions alternating every `spacing` steps
    actions = np.zeros((2, *env.action_space.shape))
    actions[1, :] = 1
    obs = env.reset()
    total = 0.0
    for i in range(30):
        obs, reward, done, info = env.step(actions[int((i % 6) >= 3)])
        total += np.mean(reward)
        assert reward.shape == (9,) and reward.dtype == np.float32
        if done:
            break
    # greedy.py: cars_on_roads().dot([1,1,-1,-1]) < 0 every `spacing` steps
    env.reset()
    for i in range(30):
        counts = env.unwrapped.cars_on_roads()
        if i % 3 == 0:
            action = env.action_space.to_action(counts.dot([1, 1, -1, -1]) < 0)
        obs, reward, done, info = env.step(action)
        if done:
            break
    # random.py
    env.reset()
    for i in range(10):
        obs, reward, done, info = env.step(env.action_space.sample())
    assert np.isfinite(obs).all()


def test_writes_to_obs_views_reach_the_device(dropin):
    gym, GridRoad, FLAGS = dropin
    np.random.seed(1)
    u = new_env(gym, GridRoad, seed=2)
    u.reset()
    u.current_phase[:] = 1
    u.elapsed[:] = 3
    u.step(np.ones(9, np.int32))   # no change of phase: elapsed keeps counting from the written value
    assert list(u.current_phase) == [1] * 9 and list(u.elapsed) == [4] * 9
