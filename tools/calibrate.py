"""Developer tool: scan road length / arrival rate for the 10x10 throughput config and report ring
occupancy, overflow rate and kernel throughput (greedy controller, auto-reset)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traffic_env_b200 import VecTrafficEnv  # noqa: E402


def run(m, n, L, lcps, E, steps, spacing=3, episode_len=120, K=10, auto_reset=True):
    env = VecTrafficEnv(m=m, n=n, length=L, num_envs=E, local_cars_per_sec=lcps, arrivals="philox", seed=2026,
                        auto_reset=auto_reset, episode_len=episode_len, ticks_per_step=K)
    env.reset()
    occ, dones, ms = [], 0, []
    s0 = env.stats()
    for s in range(steps):
        if s % spacing == 0:
            a = env.greedy_actions()
        obs, rew, done = env.step(a)
        ms.append(env.last_kernel_ms())
        dones += int(done.sum())
        if s % 10 == 9:
            c = env.cars_on_roads_flat()
            occ.append((c[:, :env.train_roads].mean(), c[:, env.train_roads:].mean(), c.max(), (c >= 14).mean()))
    s1 = env.stats()
    vu = s1["vehicle_updates"] - s0["vehicle_updates"]
    tk = s1["ticks"] - s0["ticks"]
    o = np.array(occ)
    print("ar=%d eplen=%d ticks/step %.2f " % (auto_reset, episode_len, tk / steps / E), end="")
    print("m=%d L=%g lcps=%.3f E=%d: train occ %.2f exit occ %.2f max %d frac>=14 %.3f | done/step/env %.4f | "
          "cars/env-tick %.1f | kernel %.2f ms/step -> %.3e veh-upd/s (last half), seq ticks %d" %
          (m, L, lcps, E, o[len(o)//2:, 0].mean(), o[len(o)//2:, 1].mean(), o[:, 2].max(), o[len(o)//2:, 3].mean(),
           dones / steps / E, vu / max(tk, 1), np.mean(ms[len(ms)//2:]),
           (vu / steps) / (np.mean(ms) * 1e-3), s1["seq_fallback_ticks"]))
    env.close()


if __name__ == "__main__":
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    for L in (250, 400, 500, 650):
        for lcps in (0.08, 0.10, 0.12, 0.14):
            run(10, 10, L, lcps, E, 240, episode_len=120)
    for L in (250, 500):
        for lcps in (0.12, 0.16, 0.25):
            run(10, 10, L, lcps, E, 240, auto_reset=False)
