"""Developer tool: run a raw golden fixture on the GPU and on the oracle side by side and print
the first field that differs.  Usage: python tests/debug_parity.py tests/golden/NAME.npz [max_ticks]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import OracleEnv  # noqa: E402
from tests.golden_util import unpack_schedule  # noqa: E402
from traffic_env_b200 import VecTrafficEnv  # noqa: E402


def main():
    g = np.load(sys.argv[1])
    T = int(sys.argv[2]) if len(sys.argv) > 2 else int(g["ticks"])
    ls = bool(g["learn_switch"]) if "learn_switch" in g else False
    spec = int(g["entry_spec"]) if "entry_spec" in g else 0
    env = VecTrafficEnv(m=int(g["m"]), n=int(g["n"]), length=float(g["length"]), rate=float(g["rate"]), num_envs=1,
                        arrivals="injected", learn_switch=ls, entry=spec, remi=False)
    o = OracleEnv(int(g["m"]), int(g["n"]), float(g["length"]), float(g["rate"]), learn_switch=ls)
    o.generate_entrypoints(spec)
    print("entry gpu", env.entrypoints, "oracle", o.entrypoints)
    sched = unpack_schedule(g["sched_off"], g["sched_roads"])
    env.set_arrivals([sched])
    env.reset(init_phase=g["init_phase"][None])
    o.reset(g["init_phase"])
    for t in range(T):
        obs, rew, done = env.step_raw(g["actions"][t][None])
        od = o.step(g["actions"][t], sched[t])
        st = env.get_state(0, 1)
        bad = []
        if (st["leading"][0] != o.leading).any(): bad.append(("leading", np.nonzero(st["leading"][0] != o.leading)[0][:8]))
        if (st["lastcar"][0] != o.lastcar).any(): bad.append(("lastcar", np.nonzero(st["lastcar"][0] != o.lastcar)[0][:8]))
        if (obs[0] != o.obs).any(): bad.append(("obs", np.nonzero(obs[0] != o.obs)[0][:8], obs[0][obs[0] != o.obs][:8], o.obs[obs[0] != o.obs][:8]))
        if (st["waiting"][0] != o.waiting).any(): bad.append(("waiting", np.nonzero(st["waiting"][0] != o.waiting)[0][:8]))
        if (st["passed_dst"][0] != o.passed_dst).any(): bad.append(("passed_dst",))
        if (rew[0] != o.rewards).any(): bad.append(("rewards", rew[0], o.rewards.copy()))
        if bool(done[0]) != od: bad.append(("done", done[0], od))
        if not bad:
            for e in range(o.roads):
                s = int(o.leading[e])
                while s != int(o.lastcar[e]):
                    s = 1 if s + 1 >= 20 else s + 1
                    gx, gv = st["x"][0, e, s], st["v"][0, e, s]
                    ox, ov = o.state[e, 0, s], o.state[e, 1, s]
                    if gx.tobytes() != ox.tobytes() or gv.tobytes() != ov.tobytes():
                        bad.append(("car", e, s, float(gx), float(ox), float(gv), float(ov)))
                        break
                if bad:
                    break
        if bad:
            print("tick", t, "arrivals", sched[t], "DIFF:")
            for b in bad:
                print("   ", b)
            print("gpu leading", st["leading"][0]); print("ora leading", o.leading)
            print("gpu lastcar", st["lastcar"][0]); print("ora lastcar", o.lastcar)
            return 1
    print("all", T, "ticks identical; stats", env.stats())
    return 0


if __name__ == "__main__":
    sys.exit(main())
