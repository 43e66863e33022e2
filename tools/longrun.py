"""Developer tool: long no-reset runs of the 10x10 grid to see where ring occupancy settles."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traffic_env_b200 import VecTrafficEnv

def run(L, lcps, E=256, steps=1500, spacing=3):
    env = VecTrafficEnv(m=10, n=10, length=L, num_envs=E, local_cars_per_sec=lcps, arrivals="philox", seed=2026,
                        auto_reset=False, ticks_per_step=10)
    env.reset()
    out = []
    s_prev = env.stats()
    for s in range(steps):
        if s % spacing == 0:
            a = env.greedy_actions()
        env.step(a)
        if s % 250 == 249:
            c = env.cars_on_roads_flat(); st = env.stats()
            out.append("s%d occ %.1f/%.1f t/s %.1f" % (s + 1, c[:, :400].mean(), c[:, 400:].mean(),
                       (st["ticks"] - s_prev["ticks"]) / 250 / E))
            s_prev = st
    print("L=%g lcps=%.3f: " % (L, lcps) + " | ".join(out), flush=True)
    env.close()

for L in (250, 500):
    for lcps in (0.10, 0.12, 0.14):
        run(L, lcps)
