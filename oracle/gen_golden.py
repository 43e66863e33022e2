"""Generate tests/golden/*.npz from the UNMODIFIED reference (container only).

TEST INFRASTRUCTURE.  Run:  python -m oracle.gen_golden
Needs /root/reference (read-only) - it cannot run on the GPU box, which is why
the vectors are committed.  Each fixture holds the inputs (grid, flags, initial
phases, per-tick actions, per-tick arrival road lists) and what the reference
produced (per-tick digests, done flags, rewards, periodic full checkpoints, and
for wrapped cases the Repeater/Remi observations and rewards).

Digest of one tick (see tests/golden_util.py: tick_digest):
  sha256( leading:i32 | lastcar:i32 | obs:i32 | waiting:i32 | passed_dst:u8 |
          rewards:f32 | done:u8 | live x:f32 | live v:f32 )[:8]
"live" = ring-order walk of slots in (leading, lastcar], road-major.
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, _ROOT)
from oracle import ref_harness as rh  # noqa: E402
from tests.golden_util import tick_digest, pack_schedule  # noqa: E402

OUT = os.path.join(_ROOT, "tests", "golden")


def set_flags(ref, **kw):
    F = ref.FLAGS
    base = dict(local_cars_per_sec=0.12, rate=0.5, poisson=True, entry="all", learn_switch=False, mode="train")
    base.update(kw)
    for k, v in base.items():
        setattr(F, k, v)


def run_raw(ref, name, m, n, length, ticks, action_fn, sched_seed, np_seed=0, checkpoint_every=100, **flags):
    """Bare TrafficEnv._step for `ticks` ticks; keeps stepping after overflow."""
    set_flags(ref, **flags)
    np.random.seed(np_seed)
    env = rh.make_env(ref, m, n, length, seed=sched_seed)
    sched = rh.record_schedule(ref, m, n, ticks, seed=sched_seed, entry=flags.get("entry", "all"))
    # the reference's own generator (seeded identically) must equal the recording
    env.reset()
    init_phase = env.current_phase.copy()
    I = env.graph.intersections
    actions = np.zeros((ticks, I), dtype=np.uint8)
    digests = np.zeros(ticks, dtype=np.uint64)
    dones = np.zeros(ticks, dtype=np.uint8)
    rewards = np.zeros((ticks, I), dtype=np.float32)
    gen_cars = np.zeros(ticks, dtype=np.int64)
    ck = {}
    for t in range(ticks):
        a = action_fn(t, env)
        actions[t] = np.asarray(a).astype(bool)
        obs, rew, done, _ = env.step(a)
        xs, vs = rh.live_state(env)
        digests[t] = tick_digest(env.leading, env.lastcar, env.obs, env.waiting, env.passed_dst, rew, done, xs, vs)
        dones[t] = done
        rewards[t] = rew
        gen_cars[t] = env.generated_cars
        if (t + 1) % checkpoint_every == 0 or t + 1 == ticks:
            ck["ck%d_leading" % (t + 1)] = env.leading.copy()
            ck["ck%d_lastcar" % (t + 1)] = env.lastcar.copy()
            ck["ck%d_obs" % (t + 1)] = env.obs.copy()
            ck["ck%d_waiting" % (t + 1)] = env.waiting.copy()
            ck["ck%d_x" % (t + 1)] = xs
            ck["ck%d_v" % (t + 1)] = vs
    assert int(gen_cars[-1]) == sum(len(s) for s in sched), "recorded schedule differs from generator"
    off, roads = pack_schedule(sched)
    extra = {}
    if flags.get("mode") == "validate":
        extra["trip_times"] = np.asarray(env.trip_times, dtype=np.float64)
    meta = dict(m=m, n=n, length=float(length), rate=float(ref.FLAGS.rate), ticks=ticks,
                learn_switch=int(bool(flags.get("learn_switch", False))),
                validate=int(flags.get("mode") == "validate"),
                entry_spec=0 if flags.get("entry", "all") == "all" else 0b1110)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="raw", init_phase=init_phase.astype(np.int32),
                        actions=actions, sched_off=off, sched_roads=roads, digests=digests, dones=dones,
                        rewards=rewards, generated=gen_cars, entrypoints=env.graph.entrypoints,
                        **{k: np.asarray(v) for k, v in meta.items()}, **ck, **extra)
    print("%-28s ticks=%d cars=%d overflow_ticks=%d live=%d" % (name, ticks, gen_cars[-1], dones.sum(), len(xs)))


def run_wrapped(ref, name, m, n, length, actor_steps, K, n_envs, action_seed=1234, **flags):
    """Remi(Repeater(K)(TrafficEnv)) as traffic_test.py:78-91 builds it, for
    n_envs independent envs (arrival seed = env index), Bernoulli(0.5) actions
    per intersection per actor step."""
    import traffic_test  # reference launcher: Repeater / Remi live here
    set_flags(ref, **flags)
    ref.FLAGS.light_iterations = K
    E_total = 4096
    arng = np.random.RandomState(action_seed)
    I = m * n
    init_all = arng.randint(2, size=(E_total, I)).astype(np.int32)
    act_all = arng.randint(2, size=(actor_steps, E_total, I)).astype(np.uint8)
    r = 4 * I
    obs_out = np.zeros((n_envs, actor_steps, 2 * r + I), dtype=np.float32)
    rew_out = np.zeros((n_envs, actor_steps, I), dtype=np.float32)
    done_out = np.zeros((n_envs, actor_steps), dtype=np.uint8)
    fin = {}
    offs, roads_all = [], []
    ticks = actor_steps * K
    for e in range(n_envs):
        base = rh.make_env(ref, m, n, length, seed=e)
        sched = rh.record_schedule(ref, m, n, ticks, seed=e)
        env = traffic_test.Remi(traffic_test.Repeater(K)(base))
        # TrafficEnv._reset then explicit initial phase (replaces action_space.sample());
        # Repeater._reset's extra random step is not used: the batched env exposes reset and step separately.
        base.reset()
        base.current_phase[:] = init_all[e]
        for s in range(actor_steps):
            obs, rew, done, _ = env.step(act_all[s, e].astype(np.int32))
            obs_out[e, s] = obs
            rew_out[e, s] = rew
            done_out[e, s] = done
        xs, vs = rh.live_state(base)
        fin["fin%d_leading" % e] = base.leading.copy()
        fin["fin%d_lastcar" % e] = base.lastcar.copy()
        fin["fin%d_obs" % e] = base.obs.copy()
        fin["fin%d_x" % e] = xs
        fin["fin%d_v" % e] = vs
        off, roads = pack_schedule(sched)
        offs.append(off)
        roads_all.append(roads)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="wrapped", m=m, n=n, length=float(length),
                        rate=float(ref.FLAGS.rate), K=K, actor_steps=actor_steps, n_envs=n_envs,
                        init_phase=init_all[:n_envs], actions=act_all[:, :n_envs],
                        sched_off=np.stack(offs), sched_roads=np.concatenate(roads_all),
                        sched_roads_off=np.cumsum([0] + [len(x) for x in roads_all]).astype(np.int64),
                        obs=obs_out, reward=rew_out, done=done_out, **fin)
    print("%-28s envs=%d steps=%d done_steps=%d" % (name, n_envs, actor_steps, done_out.sum()))


def run_random_entry(ref, name, episodes=4, ticks_per_episode=150):
    """FLAGS.entry == 'random' (traffic_env.py:389-394): reset_entrypoints() re-draws the open sides between
    episodes while the arrival generator (and its RandomState) carries on.  Recorded: the arrivals the reference
    actually made (its `rand.choice` results per tick), the entry sets, per-tick digests."""
    set_flags(ref, entry="random", local_cars_per_sec=0.3)
    np.random.seed(2)
    env = rh.make_env(ref, 3, 3, 250, seed=21)

    class Spy(object):
        def __init__(self, rand):
            self.rand, self.log = rand, []

        def choice(self, a):
            r = self.rand.choice(a)
            self.log.append(int(r))
            return r

        def __getattr__(self, k):
            return getattr(self.rand, k)
    spy = Spy(env.rand)
    env.rand = spy
    # NB: the generator created by seed_generator() keeps using the original RandomState object (same stream)
    digests, dones, entries, actions, per_tick, phases = [], [], [], [], [], []
    rng = np.random.RandomState(5)
    for ep in range(episodes):
        env.reset_entrypoints()
        entries.append(env.graph.entrypoints.copy())
        env.reset()
        phases.append(env.current_phase.copy())
        for t in range(ticks_per_episode):
            if t % 10 == 0:
                a = rng.randint(2, size=9).astype(np.int32)
            n0 = len(spy.log)
            obs, rew, done, _ = env.step(a)
            per_tick.append(spy.log[n0:])
            xs, vs = rh.live_state(env)
            digests.append(tick_digest(env.leading, env.lastcar, env.obs, env.waiting, env.passed_dst, rew, done, xs, vs))
            dones.append(done)
            actions.append(a.copy())
    off, roads = pack_schedule(per_tick)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="random_entry", episodes=episodes,
                        ticks_per_episode=ticks_per_episode, digests=np.asarray(digests, np.uint64),
                        dones=np.asarray(dones, np.uint8), actions=np.asarray(actions, np.uint8),
                        init_phases=np.asarray(phases, np.int32), sched_off=off, sched_roads=roads,
                        entry_sizes=np.asarray([len(e) for e in entries]), entries=np.concatenate(entries))
    print("%-28s episodes=%d entry sets=%s cars=%d" % (name, episodes, [len(e) for e in entries], len(spy.log)))
    set_flags(ref)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = rh.load()
    if len(sys.argv) > 1 and sys.argv[1] == "random_entry":
        run_random_entry(ref, "entry_random_3x3")
        return

    def fixed(t, env):  # algorithms/fixed.py:6-7 with spacing=3 at K=10
        I = env.graph.intersections
        return np.ones(I, np.int32) if ((t // 10) % 6) >= 3 else np.zeros(I, np.int32)

    def randlights(period, seed):
        rng = np.random.RandomState(seed)
        cur = {}

        def f(t, env):
            if t % period == 0:
                cur["a"] = rng.randint(2, size=env.graph.intersections).astype(np.int32)
            return cur["a"]
        return f

    # KAT-A / KAT-B of SURVEY.md 8c (same recipe; checkpoints at 200 and 1200 included)
    run_raw(ref, "kat_fixed_3x3", 3, 3, 250, 1200, fixed, sched_seed=0, checkpoint_every=200)
    # heavy traffic: ring overflow on entries and on transfers, stepping past `done`
    run_raw(ref, "overflow_3x3", 3, 3, 250, 900, randlights(40, 7), sched_seed=3, local_cars_per_sec=0.9)
    run_raw(ref, "overflow_2x2_stuck", 2, 2, 120, 500, randlights(150, 8), sched_seed=4, local_cars_per_sec=1.2)
    # validate mode: trip times of cars leaving the map (advance_hack)
    run_raw(ref, "validate_3x3", 3, 3, 250, 600, randlights(10, 9), sched_seed=5, mode="validate")
    # learn_switch, non-square grid, short roads
    run_raw(ref, "learnswitch_2x3", 2, 3, 100, 500, randlights(7, 10), sched_seed=6, learn_switch=True,
            local_cars_per_sec=0.3)
    # one entry side only; regular (non-Poisson) arrivals
    run_raw(ref, "entry_one_3x3", 3, 3, 250, 400, randlights(10, 11), sched_seed=7, entry="one",
            local_cars_per_sec=0.5)
    run_raw(ref, "regular_3x2", 3, 2, 250, 400, randlights(10, 12), sched_seed=8, poisson=False,
            local_cars_per_sec=0.25)
    # the headline grid, dense traffic
    run_raw(ref, "grid10_len500", 10, 10, 500, 400, randlights(20, 13), sched_seed=9, local_cars_per_sec=0.35)
    # config 2: Remi(Repeater(10)), 8 envs of the 4096, 120 actor steps
    run_wrapped(ref, "wrapped_3x3_cfg2", 3, 3, 250, 120, 10, 8)
    run_wrapped(ref, "wrapped_4x2_dense", 4, 2, 180, 60, 10, 4, local_cars_per_sec=0.22)
    set_flags(ref)


if __name__ == "__main__":
    main()
