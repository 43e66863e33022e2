"""Build-container tool: time the UNMODIFIED reference (numba) env on BASELINE.json configs[0] and on the headline
workload, with the C oracle port (what `bench.py --impl reference` times on the GPU box) beside it in the same run.

    python tools/time_reference.py [--seconds 8] > profiles/cpu_reference_numba.json

Needs /root/reference (read-only); the reference cannot travel to the GPU box, so this file is the record of what the
real reference does per core and how much faster the port is (the ratio that makes the GPU/CPU ratios of bench.py
conservative).

  config1   GridRoad(3,3,250), one env, `fixed` policy (spacing 3), Remi(Repeater(10)) built with the reference's own
            wrapper classes (traffic_test.py:27-64), Poisson arrivals from RandomState(seed), 120 actor steps / episode
  headline  GridRoad(10,10,500), greedy policy (greedy.py:14-16, spacing 3), same wrappers, no reset (the bench workload)

Units: a vehicle-update = one real car advanced one tick; counted by replaying the recorded actions and
arrival schedule on the oracle, which reproduces the reference's trajectory bit for bit (tests/test_reference_live.py).
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

K, SPACING, EPISODE_LEN = 10, 3, 120


def _ref_env(m, n, length, seed):
    from oracle import ref_harness as rh
    from oracle.gen_golden import set_flags
    ref = rh.load()
    set_flags(ref)
    ref.FLAGS.light_iterations = K
    import traffic_test
    base = rh.make_env(ref, m, n, length, seed=seed)
    return ref, rh, base, traffic_test.Remi(traffic_test.Repeater(K)(base))


def _numba_worker(args):
    """(config, seed, seconds) -> dict.  One reference env on one core."""
    config, seed, seconds = args
    m, n, length = (3, 3, 250) if config == "config1" else (10, 10, 500)
    ref, rh, base, env = _ref_env(m, n, length, seed)
    np.random.seed(seed)
    I = m * n
    actions, resets = [], []
    zeros, ones = np.zeros(I, np.int32), np.ones(I, np.int32)
    env.reset()                       # also compiles nothing new: JIT happened at import
    init_phase = base.current_phase.copy()
    # warm-up episode (excluded), then timed
    def run(nsteps_limit, deadline, record):
        steps = 0
        i = 0
        act = zeros
        while steps < nsteps_limit and (deadline is None or time.perf_counter() < deadline):
            if config == "config1":
                act = ones if (i % (2 * SPACING)) >= SPACING else zeros      # fixed.py:6-7
            elif i % SPACING == 0:
                act = base.action_space.to_action(base.cars_on_roads().dot([1, 1, -1, -1]) < 0)   # greedy.py:14-16
            obs, rew, done, _ = env.step(act)
            if record is not None:
                record.append(np.asarray(act, np.uint8).copy())
            steps += 1
            i += 1
            if config == "config1" and (done or i >= EPISODE_LEN):
                base.reset()          # TrafficEnv._reset (the timed quantity is the env step, not Repeater._reset's extra step)
                if record is not None:
                    resets.append((steps, base.current_phase.copy()))
                i = 0
        return steps
    run(30, None, None)
    # restart from a known state so the oracle replay can follow
    base2_ref, rh2, base, env = _ref_env(m, n, length, seed)
    np.random.seed(seed)
    base.reset()
    init_phase = base.current_phase.copy()
    t0 = time.perf_counter()
    tick0 = float(base.steps)
    steps = run(10 ** 9, t0 + seconds, actions)
    dt = time.perf_counter() - t0
    return dict(config=config, seed=seed, actor_steps=steps, seconds=dt, init_phase=init_phase.tolist(),
                actions=np.asarray(actions, np.uint8), resets=resets)


def _count_on_oracle(res, m, n, length):
    """Replay (actions, arrival schedule, resets) on the oracle -> vehicle-updates and ticks of the timed run."""
    from oracle import ref_harness as rh
    from oracle.oracle import OracleEnv
    ref = rh.load()
    acts = res["actions"]
    sched = rh.record_schedule(ref, m, n, len(acts) * K + K, seed=res["seed"])
    o = OracleEnv(m, n, float(length), 0.5)
    o.reset(np.asarray(res["init_phase"], np.int32))
    resets = dict((s, ph) for s, ph in res["resets"])
    tick = 0
    for s, a in enumerate(acts):
        for _ in range(K):
            done = o.step(a, sched[tick])
            tick += 1
            if done:
                break
        o.remi_reward()
        if (s + 1) in resets:
            o.reset(np.asarray(resets[s + 1], np.int32))
    return o.vehicle_updates, tick


def _port_worker(args):
    """The oracle port on the same workload, one env on one core: (config, seed, seconds) -> (vu, ticks, steps, dt)."""
    config, seed, seconds = args
    from oracle.oracle import OracleEnv
    from traffic_env_b200.arrivals import gap_cdf
    m, n, length = (3, 3, 250.0) if config == "config1" else (10, 10, 500.0)
    I = m * n
    o = OracleEnv(m, n, length, 0.5)
    rng = np.random.RandomState(seed)
    o.reset(rng.randint(2, size=I).astype(np.int32))
    o.philox_seed(2026, seed, gap_cdf(0.12 * m * 4 * 0.5))
    zeros, ones = np.zeros(I, np.int32), np.ones(I, np.int32)
    act = zeros
    t0 = time.perf_counter()
    vu0, tick0 = o.vehicle_updates, o.steps
    steps = i = 0
    ticks = 0.0
    while time.perf_counter() < t0 + seconds:
        if config == "config1":
            act = ones if (i % (2 * SPACING)) >= SPACING else zeros
        elif i % SPACING == 0:
            act = (o.cars_on_roads().reshape(-1, 4).dot([1, 1, -1, -1]) < 0).astype(np.int32)
        before = o.steps
        _, _, done = o.actor_step_philox(act, K, use_remi=True)
        ticks += o.steps - before
        steps += 1
        i += 1
        if config == "config1" and (done or i >= EPISODE_LEN):
            o.reset(rng.randint(2, size=I).astype(np.int32))
            i = 0
    dt = time.perf_counter() - t0
    return o.vehicle_updates - vu0, ticks, steps, dt


def measure(config, cores, seconds):
    m, n, length = (3, 3, 250) if config == "config1" else (10, 10, 500)
    ctx = mp.get_context("fork")
    jobs = [(config, s, seconds) for s in range(cores)]
    if cores == 1:
        res = [_numba_worker(jobs[0])]
        port = [_port_worker(jobs[0])]
    else:
        with ctx.Pool(cores) as pool:
            res = pool.map(_numba_worker, jobs)
        with ctx.Pool(cores) as pool:
            port = pool.map(_port_worker, jobs)
    vu = ticks = steps = 0
    for r in res:
        v, t = _count_on_oracle(r, m, n, length)
        vu += v; ticks += t; steps += r["actor_steps"]
    wall = max(r["seconds"] for r in res)
    pvu, pticks, psteps, pwall = sum(p[0] for p in port), sum(p[1] for p in port), sum(p[2] for p in port), max(p[3] for p in port)
    return {
        "cores": cores,
        "numba_reference": {"vehicle_updates_per_sec": vu / wall, "ticks_per_sec": ticks / wall,
                            "actor_steps_per_sec": steps / wall, "seconds": wall,
                            "mean_cars_per_env": vu / max(ticks, 1)},
        "c_port": {"vehicle_updates_per_sec": pvu / pwall, "ticks_per_sec": pticks / pwall,
                   "actor_steps_per_sec": psteps / pwall, "seconds": pwall, "mean_cars_per_env": pvu / max(pticks, 1),
                   "arrivals": "philox (the bench's stream), same rate as the reference's MT19937 process"},
        "port_over_numba": (pvu / pwall) / (vu / wall),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=8.0)
    a = ap.parse_args()
    from oracle import oracle as orc
    orc.build()
    from oracle import ref_harness as rh
    rh.load()                                       # JIT / cache warm-up before forking
    ncores = os.cpu_count() or 1
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = {"what": "unmodified reference (numba %s) vs the C oracle port, build container" % __import__("numba").__version__,
           "cpu": cpu, "host_cores": ncores, "seconds_per_measurement": a.seconds,
           "wrappers": "Remi(Repeater(10)) - the reference's own classes (traffic_test.py:27-64)"}
    for config in ("config1", "headline"):
        out[config] = {"grid": "3x3 L=250, fixed(spacing 3), episodes of 120 actor steps" if config == "config1"
                       else "10x10 L=500, greedy(spacing 3), no reset (bench.py grid10x10_L500_greedy, from an empty map)",
                       "one_core": measure(config, 1, a.seconds), "all_cores": measure(config, ncores, a.seconds)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
