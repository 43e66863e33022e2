"""GPU: BASELINE.json configs[1] at full size - default grid GridRoad(3,3,250), 4096 batched envs, light
actions i.i.d. Bernoulli(0.5) per intersection per actor step from RandomState(1234) (initial phases from the
same stream), injected arrival schedule of env e = the reference's own generators seeded with e, 120 actor
steps of Remi(Repeater(10)).

Checked: the 8 envs recorded from the unmodified reference (tests/golden/wrapped_3x3_cfg2.npz) match every
actor step bit for bit INSIDE the 4096-env batch, and 24 more envs spread over the batch match the oracle."""
import os

import numpy as np
import pytest

from oracle.oracle import OracleEnv
from tests.golden_util import live_walk
from tests.test_oracle_golden import oracle_actor_step
from traffic_env_b200.host_arrivals import ArrivalStream

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
E, S, K, I = 4096, 120, 10, 9


def test_config2_4096_envs():
    from traffic_env_b200 import VecTrafficEnv
    g = np.load(os.path.join(GOLDEN, "wrapped_3x3_cfg2.npz"))
    arng = np.random.RandomState(1234)
    init = arng.randint(2, size=(E, I)).astype(np.int32)
    acts = arng.randint(2, size=(S, E, I)).astype(np.uint8)
    assert (init[:8] == g["init_phase"]).all() and (acts[:, :8] == g["actions"]).all()
    env = VecTrafficEnv(m=3, n=3, length=250.0, num_envs=E, arrivals="injected", remi=True, ticks_per_step=K)
    scheds = [ArrivalStream(e, env.entrypoints, env.cars_per_sec, 0.5).window(S * K) for e in range(E)]
    env.set_arrivals(scheds)
    env.reset(init_phase=init)
    sample = sorted(set(range(8)) | set(int(v) for v in np.random.RandomState(7).randint(8, E, size=24)) | {E - 1})
    oracles, cursor = {}, {}
    for e in sample:
        o = OracleEnv(3, 3, 250.0, 0.5)
        o.reset(init[e])
        oracles[e], cursor[e] = o, 0
    done_steps = 0
    for s in range(S):
        obs, rew, done = env.step(acts[s])
        done_steps += int(done.sum())
        for e in range(8):  # the reference's own recording
            assert obs[e].tobytes() == g["obs"][e, s].tobytes(), (e, s)
            assert rew[e].tobytes() == g["reward"][e, s].tobytes() and bool(done[e]) == bool(g["done"][e, s]), (e, s)
        for e in sample:
            oo, orw, od, n = oracle_actor_step(oracles[e], acts[s, e], scheds[e], cursor[e], K)
            cursor[e] += n
            assert obs[e].tobytes() == oo.tobytes() and rew[e].tobytes() == orw.tobytes() and bool(done[e]) == od, (e, s)
    st = env.get_state()
    for e in sample:
        o = oracles[e]
        assert (st["leading"][e] == o.leading).all() and (st["lastcar"][e] == o.lastcar).all()
        gx, gv = live_walk(st["leading"][e], st["lastcar"][e], st["x"][e], st["v"][e])
        ox, ov = o.live_state()
        assert gx.tobytes() == ox.tobytes() and gv.tobytes() == ov.tobytes()
    stats = env.stats()
    assert stats["actor_steps"] == E * S and stats["ticks"] <= E * S * K
    assert done_steps > 0  # random lights do overflow some rings: the break path is exercised at scale
