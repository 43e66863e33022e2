"""Tiny workload for compute-sanitizer: a few envs, all code paths (Philox + injected arrivals, raw + fused steps,
overflow, ordered-transfer fallback, validate mode, reset, remi, greedy)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traffic_env_b200 import VecTrafficEnv

rng = np.random.RandomState(0)
env = VecTrafficEnv(m=3, n=3, num_envs=6, arrivals="philox", seed=1, local_cars_per_sec=0.9, auto_reset=True, episode_len=4)
env.reset()
for s in range(8):
    env.step(env.greedy_actions())
env2 = VecTrafficEnv(m=10, n=10, length=500.0, num_envs=3, arrivals="philox", seed=2, local_cars_per_sec=0.2)
env2.reset()
for s in range(4):
    env2.step(rng.randint(2, size=(3, 100)))
env3 = VecTrafficEnv(m=2, n=2, length=120.0, num_envs=4, arrivals="injected", remi=False, validate=True, ordered_transfers=False)
sched = [[list(rng.choice(env3.entrypoints, size=rng.randint(0, 5))) for _ in range(60)] for _ in range(4)]
env3.set_arrivals(sched)
env3.reset(init_phase=rng.randint(2, size=(4, 4)))
for t in range(60):
    env3.step_raw(rng.randint(2, size=(4, 4)))
env3.remi_reward(); env3.trip_times(); env3.cars_on_roads(); env3.get_state()
print("sanitizer case done", env.stats()["overflows"], env2.stats()["ticks"], env3.stats()["seq_fallback_ticks"])
