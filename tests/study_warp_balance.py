"""Study script (not a test; it uses the CPU oracle, hence lives under tests/): how many car-loop iterations per
tick does the slowest warp of a 10x10 env need under different road -> warp assignments?  The step kernel deals the
roads to the warps per launch (counting sort by car count, snake order); DESIGN.md section 4 quotes the numbers.

  python tests/study_warp_balance.py [warps_per_env]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import OracleEnv  # noqa: E402
from traffic_env_b200.arrivals import gap_cdf  # noqa: E402


def main():
    nw = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    m = n = 10
    R, Rp = 440, 32 * nw
    res = {"snake": [], "round_robin": [], "index_order": [], "perfect": []}
    for seed in range(6):
        o = OracleEnv(m, n, 500.0, 0.5)
        o.reset(np.zeros(m * n, np.int32))
        o.philox_seed(2026, seed, gap_cdf(0.12 * m * 4 * 0.5))
        act = np.zeros(m * n, np.int32)
        for s in range(330):
            if s % 3 == 0:
                act = (o.cars_on_roads().reshape(-1, 4).dot([1, 1, -1, -1]) < 0).astype(np.int32)
            if s < 300:
                o.actor_step_philox(act, 10, use_remi=True)
                continue
            cnt = np.zeros(Rp, int)
            cnt[:R] = o.cars_on_roads_flat()
            snake = np.zeros(Rp, int)
            for rank, rd in enumerate(np.argsort(-cnt, kind="stable")):
                row, pos = divmod(rank, nw)
                snake[rd] = (nw - 1 - pos) if row & 1 else pos
            assign = {"snake": snake, "round_robin": np.arange(Rp) % nw, "index_order": np.arange(Rp) // 32}
            for t in range(10):
                c = np.zeros(Rp, int)
                c[:R] = o.cars_on_roads_flat()
                for name, w in assign.items():
                    res[name].append(np.ceil(np.bincount(w, weights=c, minlength=nw) / 32).max())
                res["perfect"].append(np.ceil(c.sum() / (32 * nw)))
                ov0 = o.overflows
                o.actor_step_philox(act, 1, use_remi=False)
                if o.overflows > ov0:
                    break
    for k, v in res.items():
        print("%-12s mean iterations of the slowest warp per tick: %.2f" % (k, np.mean(v)))


if __name__ == "__main__":
    main()
