"""GPU: counter-based arrivals, shard invariance, auto-reset and the return statistics."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle.oracle import OracleEnv

pytestmark = pytest.mark.gpu


def make(E, base=0, **kw):
    from traffic_env_b200 import VecTrafficEnv
    args = dict(m=3, n=3, length=250.0, num_envs=E, arrivals="philox", seed=99, local_cars_per_sec=0.3,
                ticks_per_step=10, remi=True, env_id_base=base)
    args.update(kw)
    return VecTrafficEnv(**args)


def test_philox_mode_matches_oracle_actor_steps():
    """Same Philox4x32-10 stream on both sides: the device run is bit-exact against the oracle
    (including ticks cut short by overflow), for 128 envs x 60 actor steps."""
    from traffic_env_b200.arrivals import gap_cdf
    E, S, K = 128, 60, 10
    env = make(E)
    rng = np.random.RandomState(4)
    init = rng.randint(2, size=(E, 9))
    env.reset(init_phase=init)
    cdf = gap_cdf(env.cars_per_sec * 0.5)
    oracles = []
    for e in range(E):
        o = OracleEnv(3, 3, 250.0, 0.5)
        o.reset(init[e])
        o.philox_seed(99, e, cdf)
        oracles.append(o)
    n_done = 0
    for s in range(S):
        act = rng.randint(2, size=(E, 9)) if s % 3 == 0 else act
        obs, rew, done = env.step(act)
        for e, o in enumerate(oracles):
            oo, orw, od = o.actor_step_philox(act[e], K, use_remi=True)
            assert obs[e].tobytes() == oo.tobytes(), (e, s)
            assert rew[e].tobytes() == orw.tobytes(), (e, s)
            assert bool(done[e]) == od, (e, s)
            n_done += od
    st = env.stats()
    assert st["vehicle_updates"] == sum(o.vehicle_updates for o in oracles)
    assert st["cars_generated"] == sum(o.generated_cars for o in oracles)
    assert st["ticks"] == sum(int(o.steps) for o in oracles)
    assert n_done > 0, "the case should include overflow-shortened steps"


def test_shard_invariance():
    """E envs on one handle == the same global env ids split over two handles (Philox keyed by global id)."""
    E, S = 64, 25
    rng = np.random.RandomState(8)
    init = rng.randint(2, size=(E, 9))
    acts = rng.randint(2, size=(S, E, 9))
    whole = make(E)
    lo, hi = make(E // 2, base=0), make(E // 2, base=E // 2)
    whole.reset(init_phase=init)
    lo.reset(init_phase=init[:E // 2])
    hi.reset(init_phase=init[E // 2:])
    for s in range(S):
        o, r, d = whole.step(acts[s])
        o1, r1, d1 = lo.step(acts[s, :E // 2])
        o2, r2, d2 = hi.step(acts[s, E // 2:])
        assert o.tobytes() == np.concatenate([o1, o2]).tobytes()
        assert r.tobytes() == np.concatenate([r1, r2]).tobytes()
        assert (d == np.concatenate([d1, d2])).all()
    a, b, c = whole.get_state(), lo.get_state(), hi.get_state()
    for k in ("leading", "lastcar", "waiting"):
        assert (a[k] == np.concatenate([b[k], c[k]])).all()
    sa, sb, sc = whole.stats(), lo.stats(), hi.stats()
    for k in ("ticks", "vehicle_updates", "overflows", "cars_generated"):
        assert sa[k] == sb[k] + sc[k]


def reset_phases(seed, env_id, reset_count, I):
    """The device's Philox draw for the initial phases of a reset (te_reset_kernel)."""
    out = np.zeros(I, np.int32)
    for i in range(I):
        o = orc.philox4x32_10([reset_count, i >> 7, 1, 0x5e5e7], [seed, env_id])
        out[i] = (o[(i >> 5) & 3] >> (i & 31)) & 1
    return out


@pytest.mark.parametrize("m,n,L,E,S,EPL,lcps", [(3, 3, 250.0, 48, 40, 7, 0.45), (10, 10, 500.0, 160, 24, 9, 0.3),
                                                 (5, 5, 200.0, 333, 30, 6, 0.4)])
def test_auto_reset_and_return_statistics(m, n, L, E, S, EPL, lcps):
    """TE_AUTO_RESET: an env whose step overflowed, or that reached episode_len, is _reset at the start of
    its next step; episode returns accumulate as util.py:68-94 defines them.  Default grid (shared CTAs), the 10x10
    grid and a 5x5 grid (Philox reset phases of more than 32 intersections: several words of the draw)."""
    from traffic_env_b200.arrivals import gap_cdf
    K, GAMMA = 10, 0.8
    I = m * n
    env = make(E, m=m, n=n, length=L, auto_reset=True, episode_len=EPL, gamma=GAMMA, local_cars_per_sec=lcps)
    rng = np.random.RandomState(5)
    init = rng.randint(2, size=(E, I))
    env.reset(init_phase=init)
    cdf = gap_cdf(env.cars_per_sec * 0.5)
    # reset_count: 1 after te_create, 2 after the explicit reset above; it keys the Philox draw of the next reset
    oracles, ep_step, resets = [], np.zeros(E, int), 2 * np.ones(E, int)
    ret, disc, mult = np.zeros(E), np.zeros(E), np.ones(E)
    closed, ret_sum, disc_sum = 0, 0.0, 0.0
    was_done = np.zeros(E, bool)
    for e in range(E):
        o = OracleEnv(m, n, L, 0.5)
        o.reset(init[e])
        o.philox_seed(99, e, cdf)
        oracles.append(o)
    for s in range(S):
        act = rng.randint(2, size=(E, I))
        obs, rew, done = env.step(act)
        for e, o in enumerate(oracles):
            if was_done[e] or ep_step[e] >= EPL:
                closed += 1
                ret_sum += ret[e]
                disc_sum += disc[e]
                o.reset(reset_phases(99, e, int(resets[e]), I))
                resets[e] += 1
                ep_step[e] = 0
                ret[e] = disc[e] = 0.0
                mult[e] = 1.0
            oo, orw, od = o.actor_step_philox(act[e], K, use_remi=True)
            assert obs[e].tobytes() == oo.tobytes(), (e, s)
            assert rew[e].tobytes() == orw.tobytes() and bool(done[e]) == od, (e, s)
            m = float(np.mean(orw.astype(np.float64)))
            ret[e] += m
            disc[e] += mult[e] * m
            mult[e] *= GAMMA
            ep_step[e] += 1
            was_done[e] = od
    st = env.stats()
    assert st["episodes"] == closed and closed > E
    assert st["return_sum"] == pytest.approx(ret_sum, rel=1e-9, abs=1e-9)
    assert st["disc_return_sum"] == pytest.approx(disc_sum, rel=1e-6, abs=1e-9)


def test_greedy_and_cars_on_roads_match_oracle():
    E = 32
    env = make(E)
    env.reset(init_phase=np.zeros((E, 9), np.int64))
    from traffic_env_b200.arrivals import gap_cdf
    cdf = gap_cdf(env.cars_per_sec * 0.5)
    oracles = []
    for e in range(E):
        o = OracleEnv(3, 3, 250.0, 0.5)
        o.reset(np.zeros(9, np.int32))
        o.philox_seed(99, e, cdf)
        oracles.append(o)
    act = np.zeros((E, 9), np.int64)
    for s in range(12):
        c = env.cars_on_roads()
        g = env.greedy_actions()
        for e, o in enumerate(oracles):
            oc = o.cars_on_roads()
            assert (c[e] == oc).all()
            assert (g[e] == (oc.reshape(-1, 4).dot([1, 1, -1, -1]) < 0)).all()  # greedy.py:16
        act = g
        env.step(act)
        for e, o in enumerate(oracles):
            o.actor_step_philox(act[e], 10, use_remi=True)


def test_device_buffers_and_host_buffers_agree():
    """TE_DEVICE (torch CUDA tensors on a caller stream) and TE_HOST produce the same step."""
    import torch
    E = 40
    a, b = make(E), make(E)
    rng = np.random.RandomState(2)
    init = rng.randint(2, size=(E, 9))
    a.reset(init_phase=init)
    b.reset(init_phase=init)
    dev = torch.device("cuda", 0)
    d_act = torch.zeros((E, 9), dtype=torch.uint8, device=dev)
    d_obs = torch.empty((E, a.obs_len), dtype=torch.float32, device=dev)
    d_rew = torch.empty((E, 9), dtype=torch.float32, device=dev)
    d_done = torch.empty((E,), dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream()
    for s in range(10):
        act = rng.randint(2, size=(E, 9)).astype(np.uint8)
        obs, rew, done = a.step(act)
        with torch.cuda.stream(st):
            d_act.copy_(torch.from_numpy(act), non_blocking=False)
            b.step_device(d_act, d_obs, d_rew, d_done, stream=st.cuda_stream)
        st.synchronize()
        assert d_obs.cpu().numpy().tobytes() == obs.tobytes()
        assert d_rew.cpu().numpy().tobytes() == rew.tobytes()
        assert (d_done.cpu().numpy() == done).all()


def test_host_path_slices_equal_single_launch():
    """te_step with host buffers launches the batch in slices on two streams and copies each slice back while the
    next ones run; te_step with device buffers is one launch.  Same seed -> identical obs / reward / done for
    every env, including a batch size that does not divide into equal slices."""
    import torch
    from traffic_env_b200 import VecTrafficEnv
    E = 8192 + 37
    kw = dict(m=3, n=3, length=250.0, num_envs=E, arrivals="philox", seed=99, local_cars_per_sec=0.5,
              ticks_per_step=10, remi=True)
    a, b = VecTrafficEnv(**kw), VecTrafficEnv(**kw)
    rng = np.random.RandomState(5)
    init = rng.randint(2, size=(E, 9))
    a.reset(init_phase=init)
    b.reset(init_phase=init)
    dev = torch.device("cuda", 0)
    d_obs = torch.empty((E, a.obs_len), dtype=torch.float32, device=dev)
    d_rew = torch.empty((E, 9), dtype=torch.float32, device=dev)
    d_done = torch.empty((E,), dtype=torch.uint8, device=dev)
    for s in range(12):
        act = rng.randint(2, size=(E, 9)).astype(np.uint8)
        obs, rew, done = a.step(act)
        b.step_device(torch.from_numpy(act).to(dev), d_obs, d_rew, d_done)
        torch.cuda.synchronize()
        b.synchronize()
        assert obs.tobytes() == d_obs.cpu().numpy().tobytes(), s
        assert rew.tobytes() == d_rew.cpu().numpy().tobytes(), s
        assert done.tobytes() == d_done.cpu().numpy().tobytes(), s
    assert a.stats()["vehicle_updates"] == b.stats()["vehicle_updates"] > 0


@pytest.mark.parametrize("m,n,L,E,S,lcps,multi", [(3, 3, 250.0, 48, 1200, 0.12, False), (10, 10, 500.0, 4, 330, 0.12, False),
                                                   (5, 4, 150.0, 16, 400, 0.2, False), (3, 3, 250.0, 49, 900, 0.12, True),
                                                   (10, 10, 500.0, 3, 330, 0.12, True), (3, 3, 250.0, 50, 900, 0.12, 6),
                                                   (10, 10, 500.0, 3, 330, 0.12, 6), (4, 4, 120.0, 21, 360, 0.3, 6)])
def test_long_horizon_soak_vs_oracle(m, n, L, E, S, lcps, multi):
    """Long runs (up to 12000 ticks per env, no reset: the env keeps stepping after overflows, greedy lights every
    third step as in the bench) against the oracle on the same Philox stream: every actor step's observation,
    reward and done flag, and the final ring state, bit for bit.  Reaches the states a short run does not: dense
    steady state, wrapped rings everywhere, the rare full-powf and generic-arithmetic lanes.  multi = True: one launch per
    greedy decision (3 actor steps); multi = 6: two decisions per launch (te_set_controller_spacing) - the second one is
    taken inside the kernel and checked here against the oracle's ring counts at that moment."""
    from traffic_env_b200 import VecTrafficEnv
    from traffic_env_b200.arrivals import gap_cdf
    from tests.golden_util import live_walk
    K, I = 10, m * n
    env = VecTrafficEnv(m=m, n=n, length=L, num_envs=E, arrivals="philox", seed=7, local_cars_per_sec=lcps,
                        ticks_per_step=K, remi=True)
    init = np.random.RandomState(2).randint(2, size=(E, I))
    env.reset(init_phase=init)
    cdf = gap_cdf(env.cars_per_sec * 0.5)
    oracles = []
    for e in range(E):
        o = OracleEnv(m, n, L, 0.5)
        o.reset(init[e])
        o.philox_seed(7, e, cdf)
        oracles.append(o)
    act = np.zeros((E, I), np.uint8)
    for s in range(S):
        if s % 3 == 0:
            if multi == 6:   # the bench's scheme: six actor steps = two greedy decisions per launch, controller in the kernel
                if s % 6 == 0:
                    acts6, obs6, rew6, done6 = env.step_multi(6, controller="greedy", spacing=3)
                    acts6 = acts6.copy()
                half = (s % 6) // 3
                act, obs3, rew3, done3 = acts6[half], obs6[3 * half:], rew6[3 * half:], done6[3 * half:]
            elif multi:   # the three actor steps of a greedy decision are one launch
                act, obs3, rew3, done3 = env.step_multi(3, controller="greedy")
                act = act.copy()
            else:
                act = env.greedy_actions().copy()
            for e, o in enumerate(oracles):   # the device controller == greedy.py:14-16 on the oracle's counts
                assert (act[e] == (o.cars_on_roads().reshape(-1, 4).dot([1, 1, -1, -1]) < 0)).all(), (e, s)
        obs, rew, done = (obs3[s % 3], rew3[s % 3], done3[s % 3]) if multi else env.step(act)
        for e, o in enumerate(oracles):
            oo, orw, od = o.actor_step_philox(act[e].astype(np.int32), K, use_remi=True)
            assert obs[e].tobytes() == oo.tobytes() and rew[e].tobytes() == orw.tobytes() and bool(done[e]) == od, (e, s)
    st = env.get_state()
    for e, o in enumerate(oracles):
        assert (st["leading"][e] == o.leading).all() and (st["lastcar"][e] == o.lastcar).all()
        gx, gv = live_walk(st["leading"][e], st["lastcar"][e], st["x"][e], st["v"][e])
        ox, ov = o.live_state()
        assert gx.tobytes() == ox.tobytes() and gv.tobytes() == ov.tobytes()
    assert env.stats()["vehicle_updates"] == sum(o.vehicle_updates for o in oracles)
