"""Developer tool (GPU box): where does the host path's time go?  Per actor step of the headline workload through
VecTrafficEnv.step (host buffers): wall time, GPU-side span (first slice launch -> last copy, CUDA events), the greedy
controller call, and the device-buffer launch for comparison; swept over TE_HOST_SLICES x TE_HOST_THREADS.

  python tools/e2e_probe.py            # sweep (spawns itself per setting)
"""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one():
    import torch
    from traffic_env_b200 import VecTrafficEnv
    wl = os.environ.get("PROBE_WL", "10x10")
    if wl == "10x10":
        E, kw = 16384, dict(m=10, n=10, length=500.0)
        pre = 300
    else:
        E, kw = 131072, dict(m=3, n=3, length=250.0)
        pre = 150
    env = VecTrafficEnv(num_envs=E, local_cars_per_sec=0.12, arrivals="philox", seed=2026, ticks_per_step=10, remi=True, **kw)
    env.reset()
    I = env.intersections
    h_act = np.zeros((E, I), np.uint8)
    d_act = torch.zeros((E, I), dtype=torch.uint8, device='cuda'); d_obs = torch.empty((E, env.obs_len), dtype=torch.float32, device='cuda')
    d_rew = torch.empty((E, I), dtype=torch.float32, device='cuda'); d_done = torch.empty((E,), dtype=torch.uint8, device='cuda')
    for s in range(pre):
        if s % 3 == 0:
            env.greedy_actions(out=d_act)
        env.step_device(d_act, d_obs, d_rew, d_done)
    torch.cuda.synchronize()
    h_act[:] = d_act.cpu().numpy()

    def timeit(f, n=30):
        for i in range(3):
            f(i)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(n):
            f(i)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
    t_step = timeit(lambda i: env.step(h_act))
    span = env.last_kernel_ms()
    t_wire = timeit(lambda i: env.step_wire(h_act))
    t_greedy = timeit(lambda i: env.greedy_actions())

    def hs(i):
        if i % 3 == 0:
            h_act[:] = env.greedy_actions()
        o, r, d = env.step(h_act); return float(r[0, 0]) + float(o[0, 0])
    t_bench = timeit(hs)
    t_dev = timeit(lambda i: env.step_device(d_act, d_obs, d_rew, d_done))
    print("%s slices=%s threads=%s: step %.3f ms (gpu span %.3f), step_wire %.3f, greedy_actions %.3f, bench host_step %.3f, "
          "device step %.3f" % (wl, os.environ.get("TE_HOST_SLICES", "dflt"), os.environ.get("TE_HOST_THREADS", "dflt"),
                                t_step, span, t_wire, t_greedy, t_bench, t_dev), flush=True)


if __name__ == "__main__":
    if os.environ.get("PROBE_CHILD"):
        one()
    else:
        for wl in ("10x10", "3x3"):
            for sl, th in (("", ""), ("32", "1"), ("32", "2"), ("32", "8"), ("16", "4"), ("64", "4"), ("8", "4")):
                env = dict(os.environ, PROBE_CHILD="1", PROBE_WL=wl)
                if sl:
                    env["TE_HOST_SLICES"], env["TE_HOST_THREADS"] = sl, th
                subprocess.run([sys.executable, os.path.abspath(__file__)], env=env)
