"""Config 5 (BASELINE.json configs[4]): rollout loops with the learners' contract
(algorithms/a3c.py:52-63 `epoch`: obs -> policy -> env.step -> (obs, reward, done)) against the batched GPU env,
with a stand-in policy of the learners' I/O shape (obs f32[81] -> bool[9]; TensorFlow 1 cannot run here).

  mode "threads": T Python threads (FLAGS.threads), each on its own single-env proxy (traffic_env_b200.pool); the
                  device steps whichever slots have an action queued with one masked launch (no lock-step round: a
                  slow thread only delays itself); host obs / action traffic every step.
  mode "device" : one loop over E env slots, policy = one torch matmul on the device, obs / reward / done stay in
                  HBM (VecTrafficEnv.step_device), auto-reset on done or after 120 actor steps (episode_len).

Prints one JSON line per mode: actor-steps/s (env slots x steps / wall time) and vehicle-updates/s.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_threads(T, steps):
    from traffic_env_b200.pool import EnvPool
    pool = EnvPool(T, m=3, n=3, length=250.0, arrivals="philox", seed=3, local_cars_per_sec=0.12, ticks_per_step=10)

    def worker(env, seed):
        rng = np.random.RandomState(seed)
        w = rng.standard_normal((env.observation_space.size, env.action_space.size)).astype(np.float32)
        obs = env.reset()
        for _ in range(steps):
            obs, reward, done, _ = env.step((obs @ w) < 0)
            if done:
                obs = env.reset()
        env.close()

    ths = [threading.Thread(target=worker, args=(pool.slot(i), i)) for i in range(T)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    st = pool.vec.stats()
    return {"mode": "threads", "threads": T, "steps_per_thread": steps, "actor_steps_per_sec": st["actor_steps"] / dt,
            "vehicle_updates_per_sec": st["vehicle_updates"] / dt, "launches": pool.launches,
            "mean_slots_per_launch": pool.stepped / max(pool.launches, 1), "wall_s": dt}


def run_device(E, steps):
    import torch
    from traffic_env_b200 import VecTrafficEnv
    dev = torch.device("cuda", 0)
    env = VecTrafficEnv(m=3, n=3, length=250.0, num_envs=E, arrivals="philox", seed=3, local_cars_per_sec=0.12,
                        ticks_per_step=10, remi=True, auto_reset=True, episode_len=120)
    env.reset()
    I, OL = env.intersections, env.obs_len
    g = torch.Generator(device=dev).manual_seed(0)
    w = torch.randn((OL, I), device=dev, generator=g)
    obs = torch.zeros((E, OL), dtype=torch.float32, device=dev)
    rew = torch.empty((E, I), dtype=torch.float32, device=dev)
    done = torch.empty((E,), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        act = ((obs @ w) < 0).to(torch.uint8)           # the policy: one matmul, actions stay on the device
        env.step_device(act, obs, rew, done, stream=stream)

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    s0 = env.stats()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    s1 = env.stats()
    return {"mode": "device", "env_slots": E, "steps": steps,
            "actor_steps_per_sec": (s1["actor_steps"] - s0["actor_steps"]) / dt,
            "vehicle_updates_per_sec": (s1["vehicle_updates"] - s0["vehicle_updates"]) / dt,
            "episodes": s1["episodes"] - s0["episodes"], "mean_return": (s1["return_sum"] - s0["return_sum"]) / max(1, s1["episodes"] - s0["episodes"]),
            "wall_s": dt}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--thread-steps", type=int, default=240)
    ap.add_argument("--envs", type=int, default=131072)
    ap.add_argument("--steps", type=int, default=240)
    a = ap.parse_args()
    print(json.dumps(run_threads(a.threads, a.thread_steps)), flush=True)
    print(json.dumps(run_device(a.envs, a.steps)), flush=True)
