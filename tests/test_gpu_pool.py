"""GPU: config 5 - learner-style rollout loops (a3c.epoch contract, a3c.py:52-63) from several Python
threads, each on its own single-env proxy, while the device steps all slots in one launch per round."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rollout(env, policy_seed, steps, out):
    """a3c.epoch restated: obs = env.reset(); y = policy(obs); new_obs, reward, done, _ = env.step(y)."""
    rng = np.random.RandomState(policy_seed)
    w = rng.standard_normal((env.observation_space.size, env.action_space.size)).astype(np.float32)
    obs = env.reset()
    traj = []
    for t in range(steps):
        y = (obs @ w) < 0                      # bool vector, like tf.less in algorithms/util.py:114
        new_obs, reward, done, _ = env.step(y)
        traj.append((obs.copy(), y.copy(), reward.copy(), done))
        obs = new_obs
        if done:
            obs = env.reset()
    out.append(traj)
    env.close()


def test_threads_share_one_launch_per_round():
    from traffic_env_b200.pool import EnvPool
    T, S = 4, 25
    np.random.seed(0)
    pool = EnvPool(T, m=3, n=3, length=250.0, arrivals="philox", seed=3, local_cars_per_sec=0.12, ticks_per_step=10)
    outs = [[] for _ in range(T)]
    threads = [threading.Thread(target=rollout, args=(pool.slot(i), 10 + i, S, outs[i])) for i in range(T)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
        assert not th.is_alive()
    for i in range(T):
        traj = outs[i][0]
        assert len(traj) == S
        obs, y, rew, done = traj[-1]
        assert obs.shape == (81,) and obs.dtype == np.float32 and rew.shape == (9,) and y.dtype == np.bool_
    # every round stepped all active slots at once: far fewer launches than T * S single-env steps
    assert pool.launches <= S + 2 * T + 4
    assert pool.vec.stats()["actor_steps"] >= T * S


def test_single_slot_matches_batched_env():
    """A lone slot behaves like the batched env driven directly (same seeds, same actions)."""
    from traffic_env_b200 import VecTrafficEnv
    from traffic_env_b200.pool import EnvPool
    kw = dict(m=3, n=3, length=250.0, arrivals="philox", seed=9, local_cars_per_sec=0.2, ticks_per_step=10)
    pool = EnvPool(1, **kw)
    vec = VecTrafficEnv(num_envs=1, remi=True, **kw)
    np.random.seed(5)
    env = pool.slot(0)
    o1 = env.reset()
    np.random.seed(5)
    phases = np.random.randint(2, size=(1, 9))
    vec.reset(mask=np.ones(1), init_phase=phases)
    first = np.random.randint(np.int32(2), size=[9], dtype=np.int32)
    o2 = vec.step(first[None])[0][0].copy()
    assert o1.tobytes() == o2.tobytes()
    rng = np.random.RandomState(1)
    for _ in range(15):
        a = rng.randint(2, size=9)
        x1, r1, d1, _ = env.step(a)
        x2, r2, d2 = vec.step(a[None])
        assert x1.tobytes() == x2[0].tobytes() and r1.tobytes() == r2[0].tobytes() and d1 == bool(d2[0])
