"""TEST SUPPORT - launcher for the reference's baseline controllers on the B200 env:

    python -m tests.support.run_baselines --trainer fixed|random|greedy|spacedgreedy|const0|const1 [--episodes N] [...]

Mirrors `python traffic_test.py --trainer X` (traffic_test.py:93-95, alg_flags.py:46-49) for the TensorFlow-free
controllers (algorithms/{fixed,random,greedy,spacedgreedy,const0,const1}.py): same env factory, same episode loop,
same per-episode return (util.py:68-94: sum of mean-over-intersections reward, gamma-discounted when
--print_discounted) and running mean / std print-out (util.py:13-34).  With the reference checkout available the
unmodified launcher works too: see INTEGRATION.md.
"""
import argparse
import math
import sys

import numpy as np


def controllers(env, spacing):
    shape = env.action_space.shape

    def fixed(i, state):
        return np.ones(shape) if (i % (2 * spacing)) >= spacing else np.zeros(shape)        # fixed.py:6-7,13-17

    def greedy(i, state):
        if i % spacing == 0:                                                               # greedy.py:14-16
            state["a"] = env.action_space.to_action(env.unwrapped.cars_on_roads().dot([1, 1, -1, -1]) < 0)
        return state["a"]

    return {"fixed": fixed, "random": lambda i, s: env.action_space.sample(), "greedy": greedy, "spacedgreedy": greedy,
            "const0": lambda i, s: np.zeros(shape), "const1": lambda i, s: np.ones(shape)}


def run_episode(env, policy, episode_len, gamma, discounted):
    env.reset()
    total, mult, state = 0.0, 1.0, {"a": np.zeros(env.action_space.shape)}
    for i in range(episode_len):
        obs, reward, done, info = env.step(policy(i, state))
        total += float(np.mean(reward)) * (mult if discounted else 1.0)
        mult *= gamma
        if done:
            break
    return total, i + 1


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--trainer", default="fixed", choices=["fixed", "random", "greedy", "spacedgreedy", "const0", "const1"])
    ap.add_argument("--episodes", type=int, default=10)
    ap.add_argument("--episode_secs", type=int, default=600)
    ap.add_argument("--light_secs", type=int, default=5)
    ap.add_argument("--rate", type=float, default=0.5)
    ap.add_argument("--spacing", type=int, default=3)
    ap.add_argument("--gamma", type=float, default=0.8)
    ap.add_argument("--print_discounted", type=int, default=1)
    ap.add_argument("--local_cars_per_sec", type=float, default=0.12)
    ap.add_argument("--grid", default="3x3x250")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--remi", type=int, default=1)
    a = ap.parse_args(argv)
    from tests.support import install_dropin
    install_dropin()
    from args import FLAGS
    from tests.support.wrappers_ref import make_env
    FLAGS.rate, FLAGS.local_cars_per_sec, FLAGS.poisson, FLAGS.entry, FLAGS.learn_switch = a.rate, a.local_cars_per_sec, True, "all", False
    m, n, length = (int(v) for v in a.grid.split("x"))
    np.random.seed(a.seed)
    env = make_env(m, n, length, seed=a.seed, light_iterations=int(a.light_secs / a.rate), remi=bool(a.remi))
    policy = controllers(env, a.spacing)[a.trainer]
    episode_len = int(a.episode_secs / a.light_secs)
    mean = var = 0.0
    import time
    t0, total_steps = time.perf_counter(), 0
    for it in range(1, a.episodes + 1):
        reward, steps = run_episode(env, policy, episode_len, a.gamma, bool(a.print_discounted))
        if it == 1:   # the first episode creates the CUDA context and the device handle: keep it out of the rate
            t0, total_steps = time.perf_counter(), 0
        else:
            total_steps += steps
        mean = (reward + (it - 1) * mean) / it
        if it >= 2:
            var = (it - 2) / (it - 1) * var + (reward - mean) ** 2 / it
        print("Reward %2f\t Mean %2f\t Std %2f\t (%d actor steps)" % (reward, mean, math.sqrt(var), steps))
    dt = time.perf_counter() - t0
    k = int(a.light_secs / a.rate)
    if total_steps:
        print("%d actor steps in %.2f s after the first episode: %.0f actor steps/s, up to %.0f physics ticks/s "
              "(single env, one launch per actor step)" % (total_steps, dt, total_steps / dt, total_steps * k / dt))
    return mean


if __name__ == "__main__":
    main()
