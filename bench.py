"""bench.py - vehicle-updates/s of the traffic-env simulation step on B200.

Workload (BASELINE.json configs[2], the config the target is quoted on): 10x10 grid, long
roads (L = 500 m), 16384 env instances per GPU, greedy light controller recomputed every 3
actor steps (algorithms/greedy.py:13-16 with --spacing 3), Philox arrivals at the reference's
stock --local_cars_per_sec 0.12, Remi(Repeater(10)) semantics, envs pre-rolled to the
ring-capacity-bound steady state (see DESIGN.md "headline workload").  A step = one actor step
(10 physics ticks unless a ring overflows) of every env = one te_step kernel launch.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (one rank per GPU under torchrun)
  python bench.py --impl reference ...                          the CPU arm: the oracle port on the host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (m, n, length, local_cars_per_sec, default envs per GPU, pre-roll actor steps)
    "grid10x10_L500_greedy": dict(m=10, n=10, length=500.0, lcps=0.12, envs=16384, preroll=300),
    "grid3x3_L250_greedy": dict(m=3, n=3, length=250.0, lcps=0.12, envs=131072, preroll=150),
}
K_TICKS = 10      # FLAGS.light_iterations = light_secs / rate = 5 / 0.5 (traffic_test.py:21)
SPACING = 3       # FLAGS.spacing (alg_flags.py:22)
OPS_PER_UPDATE = 34  # arithmetic ops of one sim() element


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="grid10x10_L500_greedy", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="env instances per GPU (default: the workload's)")
    ap.add_argument("--preroll", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(args):
    """One env of the workload on one core: pre-roll, then timed actor steps.  Returns (vehicle_updates, seconds,
    env_ticks, actor_steps)."""
    wl, seed, preroll, budget_s, min_steps = args
    from oracle.oracle import OracleEnv
    from traffic_env_b200.arrivals import gap_cdf
    w = WORKLOADS[wl]
    o = OracleEnv(w["m"], w["n"], w["length"], 0.5)
    o.reset(np.zeros(w["m"] * w["n"], np.int32))
    cps = w["lcps"] * w["m"] * 4
    o.philox_seed(2026, seed, gap_cdf(cps * 0.5))
    act = np.zeros(w["m"] * w["n"], np.int32)

    def run(nsteps, deadline=None):
        nonlocal act
        done_steps = 0
        for s in range(nsteps):
            if s % SPACING == 0:
                act = (o.cars_on_roads().reshape(-1, 4).dot([1, 1, -1, -1]) < 0).astype(np.int32)
            o.actor_step_philox(act, K_TICKS, use_remi=True)
            done_steps += 1
            if deadline is not None and done_steps >= min_steps and time.perf_counter() > deadline:
                break
        return done_steps

    run(preroll)
    vu0, t0 = o.vehicle_updates, time.perf_counter()
    tick0 = o.steps
    n = run(10 ** 9, deadline=t0 + budget_s)
    dt = time.perf_counter() - t0
    return o.vehicle_updates - vu0, dt, float(o.steps - tick0), n


def cpu_run(wl, cores, preroll, budget_s, min_steps=3):
    import multiprocessing as mp
    jobs = [(wl, i, preroll, budget_s, min_steps) for i in range(cores)]
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    vu = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return dict(vehicle_updates=vu, seconds=wall, value=vu / wall, env_ticks=sum(r[2] for r in res),
                actor_steps=sum(r[3] for r in res))


def reference_arm(a):
    """The reference's CPU implementation of the path, timed on the host cores.  The reference is Python +
    numba and cannot travel to the GPU box, so this is its C restatement (oracle/, kind "port"), one
    independent env per core (the reference has no intra-env parallelism)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    cores = os.cpu_count() or 1
    w = WORKLOADS[a.workload]
    per_step = []
    preroll = 60  # bounded: enough for the first cars to cross several intersections
    cpu_run(a.workload, 1, 2, 0.01, 1)  # warm the page cache / library load
    t0 = time.perf_counter()
    tot_vu, tot_s, tot_ticks, tot_steps = 0, 0.0, 0.0, 0
    budget = max(1.0, min(20.0, 150.0 / max(1, a.steps + a.warmup)))
    for s in range(a.warmup + a.steps):
        r = cpu_run(a.workload, cores, preroll, budget)
        if s >= a.warmup:
            tot_vu += r["vehicle_updates"]; tot_s += r["seconds"]; tot_ticks += r["env_ticks"]; tot_steps += r["actor_steps"]
            per_step.append(r["seconds"])
    value = tot_vu / tot_s
    sample = ("%d independent envs (one per core) of %s, %d-actor-step pre-roll then ~%.1f s of actor steps per bench step"
              % (cores, a.workload, preroll, budget))
    line = {
        "impl": "reference", "metric": "vehicle_updates_per_sec", "value": value, "unit": "vehicle-updates/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot_s / max(1, a.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": bench_config(a, w, w["envs"] if not a.envs else a.envs, None),
        "env_steps_per_sec": {"ticks": tot_ticks / tot_s, "actor_steps": tot_steps / tot_s},
        "cpu_baseline": {"value": value, "unit": "vehicle-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "vehicle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def bench_config(a, w, envs, occ, rows_padded=None):
    return {"workload": a.workload, "grid": "%dx%d" % (w["m"], w["n"]), "road_length_m": w["length"],
            "envs_per_gpu": envs, "policy": "greedy(spacing=%d)" % SPACING, "ticks_per_actor_step": K_TICKS,
            "arrivals": "philox, local_cars_per_sec=%.2f (reference default)" % w["lcps"],
            "wrappers": "Remi(Repeater(10))", "reset": "none (env keeps stepping after overflow, as the reference env does)",
            "steady_state_cars_per_env": occ, "l2": "state is %.2f GB per GPU, far larger than the 126 MB L2"
            % (envs * (rows_padded or (w["m"] * w["n"] * 4 + 2 * w["m"] + 2 * w["n"] + 31) // 32 * 32) * 160 / 1e9)}


# --------------------------------------------------------------------------- our arm
def b200_arm(a):
    import torch
    import torch.distributed as dist
    from traffic_env_b200 import VecTrafficEnv
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL prints while the communicator comes up (its
        # "NCCL version ..." banner at NCCL_DEBUG=VERSION/WARN) is sent to stderr by pointing fd 1 at fd 2 meanwhile
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    w = WORKLOADS[a.workload]
    E = a.envs or w["envs"]
    preroll = w["preroll"] if a.preroll < 0 else a.preroll
    env = VecTrafficEnv(m=w["m"], n=w["n"], length=w["length"], num_envs=E, local_cars_per_sec=w["lcps"],
                        arrivals="philox", seed=2026, env_id_base=rank * E, device=local, ticks_per_step=K_TICKS,
                        remi=True, auto_reset=False)
    I, OL = env.intersections, env.obs_len
    d_act = torch.zeros((E, I), dtype=torch.uint8, device=dev)
    d_obs = torch.empty((E, OL), dtype=torch.float32, device=dev)
    d_rew = torch.empty((E, I), dtype=torch.float32, device=dev)
    d_done = torch.empty((E,), dtype=torch.uint8, device=dev)
    # a dedicated (non-default) stream: the kernels are launched on it and the CUDA events are recorded on it
    tstream = torch.cuda.Stream(device=dev)
    stream = tstream.cuda_stream
    env.reset()
    torch.cuda.synchronize()
    launches = [0]

    def device_step(s):
        if s % SPACING == 0:
            env.greedy_actions(out=d_act, stream=stream)
            launches[0] += 1
        env.step_device(d_act, d_obs, d_rew, d_done, stream=stream)
        launches[0] += 1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # nvidia-smi needs a few hundred ms to deliver its first sample; the pre-roll, the warm-up
                          # and the timed region run the same kernel back to back, so all samples are under load
    for s in range(preroll):
        device_step(s)
    torch.cuda.synchronize()
    occ = float(env.cars_on_roads_flat().sum(axis=1).mean())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(a.warmup):
        device_step(s)
    barrier()
    st0 = env.stats()
    launches[0] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(tstream)
    for s in range(a.steps):
        device_step(s)
    ev1.record(tstream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    st1 = env.stats()
    n_launch = launches[0]
    vu = st1["vehicle_updates"] - st0["vehicle_updates"]
    ticks = st1["ticks"] - st0["ticks"]
    asteps = st1["actor_steps"] - st0["actor_steps"]
    gen = st1["cars_generated"] - st0["cars_generated"]
    seqfb = st1["seq_fallback_ticks"] - st0["seq_fallback_ticks"]
    # episode-return style reduction over NCCL/NVLink: a handful of scalars, the only collective of the path
    red = torch.tensor([float(vu), float(ticks), float(asteps), float(gen), float(st1["overflows"] - st0["overflows"])],
                       dtype=torch.float64, device=dev)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    vu_all, ticks_all, asteps_all, gen_all, ovf_all = [float(x) for x in red.tolist()]
    ms_max = float(tmax.item())

    # per-launch kernel time (CUDA events on the launch stream, separate pass so the sync does not sit in the timed region)
    kms, kvu = [], []
    for s in range(min(a.steps, 10)):
        b0 = env.stats()["vehicle_updates"]
        device_step(s)
        kms.append(env.last_kernel_ms())
        kvu.append(env.stats()["vehicle_updates"] - b0)
    k_ms = float(np.mean(kms))
    cars_env = float(np.mean(kvu)) / E / (ticks / max(asteps, 1))
    R, r = env.roads, env.train_roads
    arr_per_step = gen / max(asteps, 1)
    bytes_env = 2 * (8 * cars_env + 8 * R) + 4 * I + 4 * (2 * r + I) + 4 * I + 2 * 8 * I + 1 + 2 * arr_per_step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_env * E / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_bytes.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(a.workload)
        except Exception:
            traffic = None

    # end to end through the public API with HOST buffers: what the greedy agent does (greedy.py:13-17)
    e2e = None
    if not a.no_e2e:
        h_act = np.zeros((E, I), dtype=np.uint8)

        def host_step(s):
            # greedy agent loop (greedy.py:13-17) on the batched env through the public host API: every `spacing`
            # steps the controller output comes back to the host (te_greedy_actions, D2H); every step the actions
            # go host -> device and obs / reward / done come device -> host.
            if s % SPACING == 0:
                h_act[:] = env.greedy_actions()
            obs, rew, done = env.step(h_act)
            return float(rew[0, 0]) + float(obs[0, 0])  # the result is read on the host

        for s in range(max(3, min(a.warmup, 5))):
            host_step(s)
        barrier()
        b0 = env.stats()["vehicle_updates"]
        t0 = time.perf_counter()
        ne = max(3, a.steps // 2)
        for s in range(ne):
            host_step(s)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e_vu = torch.tensor([float(env.stats()["vehicle_updates"] - b0)], dtype=torch.float64, device=dev)
        e_t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_vu, op=dist.ReduceOp.SUM)
            dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": float(e_vu.item()) / float(e_t.item()), "unit": "vehicle-updates/s",
               "h2d_bytes_per_step": int(E * I), "d2h_bytes_per_step": int(E * (OL * 4 + I * 4 + 1) + E * I // SPACING),
               "steps": ne, "api": "VecTrafficEnv.step(actions) -> te_step(TE_HOST): actions H2D, obs/reward/done D2H into the env's "
               "page-locked host buffers every step; VecTrafficEnv.greedy_actions() (device controller, actions D2H) every %d steps" % SPACING}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        r1 = cpu_run(a.workload, 1, 60, a.cpu_seconds)
        cpu = {"value": r1["value"], "unit": "vehicle-updates/s", "cores": 1, "kind": "port",
               "sample": "1 env of %s on 1 core (oracle/traffic_oracle.c): 60-actor-step pre-roll, then %.0f s of "
               "actor steps (%d steps)" % (a.workload, r1["seconds"], r1["actor_steps"])}

    barrier()
    flush_gbs = env.stage_bandwidth(5) if rank == 0 else None
    arith_peak = None
    if rank == 0:
        from traffic_env_b200.vec_env import idm_arithmetic_peak
        arith_peak = idm_arithmetic_peak(device=local)
    if rank == 0:
        sec = ms_max * 1e-3
        value = vu_all / sec
        fp32_peak = 148 * 128 * 2 * (clocks["sm_mhz"] or 1965.0) * 1e6 if clocks else None
        line = {
            "metric": "vehicle_updates_per_sec", "value": value, "unit": "vehicle-updates/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": bench_config(a, w, E, occ, env.roads_padded),
            "env_steps_per_sec": {"ticks": ticks_all / sec, "actor_steps": asteps_all / sec},
            "ticks_per_actor_step": ticks_all / max(asteps_all, 1.0), "overflows_per_actor_step": ovf_all / max(asteps_all, 1.0),
            "ordered_transfer_ticks_frac_rank0": seqfb / max(ticks, 1),
            "target_8gpu": 1e11, "frac_of_per_gpu_target": value / world / 1.25e10,
            "clocks": clocks, "e2e": e2e, "gpu_launches": n_launch,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "te_step_kernel", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_env_step": bytes_env, "cars_per_env": cars_env,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "note": "the fused K-tick kernel is issue/FP64-pipe bound, not HBM bound (SURVEY.md 8d); see roofline_issue"},
            "roofline_flush": {"bound": "hbm", "achieved": flush_gbs, "peak": peak, "unit": "GB/s", "frac": flush_gbs / peak,
                               "what": "the step kernel's bulk-TMA stage-in + flush alone (te_stage_kernel: same CTA shape "
                               "and shared-memory footprint, no ticks), read + written bytes"},
            "roofline_compute": {"bound": "idm arithmetic (registers only, all lanes busy: te_idm_peak micro-kernel)",
                                 "achieved": float(np.mean(kvu)) / (k_ms * 1e-3), "peak": arith_peak,
                                 "unit": "vehicle-updates/s", "frac": float(np.mean(kvu)) / (k_ms * 1e-3) / arith_peak},
            "roofline_issue": {"ops_per_vehicle_update": OPS_PER_UPDATE,
                               "achieved_gops": float(np.mean(kvu)) * OPS_PER_UPDATE / (k_ms * 1e-3) / 1e9,
                               "fp32_peak_gops_nominal_at_clock": fp32_peak / 1e9 if fp32_peak else None},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        b200_arm(a)


if __name__ == "__main__":
    main()
