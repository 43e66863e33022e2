"""CPU study (oracle-driven): shared-memory bank conflicts of the car loop's x/v accesses as a function of the row pitch.

In the car loop lane l of a warp handles cars [l q, l q + q) of the warp's car list; in iteration i the 32 lanes touch
32 different (row, slot) pairs.  Word address = row * pitch + slot, bank = address mod 32.  A request needs as many
wavefronts as the fullest bank has distinct words.  This script replays the headline workload on the oracle, rebuilds
the kernel's road -> (warp, lane) assignment (counting sort by car count, snake deal) every actor step and counts
wavefronts per request for several pitches and for an ideal hash (uniformly random banks).

  python tools/bank_conflict_study.py            # ~1 min
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import OracleEnv  # noqa: E402
from traffic_env_b200.arrivals import gap_cdf  # noqa: E402

M, N, L, K, NWARPS = 10, 10, 500.0, 10, 16
PITCHES = (20, 21, 24, 33)


def main():
    rng = np.random.RandomState(0)
    o = OracleEnv(M, N, L, 0.5)
    o.reset(np.zeros(M * N, np.int32))
    o.philox_seed(2026, 0, gap_cdf(0.12 * M * 4 * 0.5))
    R = o.roads
    Rp = NWARPS * 32
    act = np.zeros(M * N, np.int32)
    req = 0
    waves = {p: 0 for p in PITCHES}
    waves["random"] = 0
    for s in range(330):
        if s % 3 == 0:
            act = (o.cars_on_roads().reshape(-1, 4).dot([1, 1, -1, -1]) < 0).astype(np.int32)
        if s >= 300:
            # assignment at launch: rows sorted by car count (descending, stable by index), dealt in snake order
            n0 = np.zeros(Rp, np.int64)
            n0[:R] = o.cars_on_roads_flat()
            order = np.argsort(-n0, kind="stable")
            warps = [[] for _ in range(NWARPS)]
            for rank, row in enumerate(order):
                r_, pos = divmod(rank, NWARPS)
                warps[NWARPS - 1 - pos if r_ & 1 else pos].append(row)
            ld, lc = o.leading.copy(), o.lastcar.copy()
            for w in warps:
                cars = []          # (row, slot) in list order
                for row in w:
                    if row >= R:
                        continue
                    sl = int(ld[row])
                    while sl != int(lc[row]):
                        sl = 1 if sl + 1 >= 20 else sl + 1
                        cars.append((row, sl))
                total = len(cars)
                if not total:
                    continue
                q = (total + 31) // 32
                for i in range(q):
                    acc = [cars[l * q + i] for l in range(32) if l * q + i < min(total, l * q + q)]
                    if not acc:
                        continue
                    req += 1
                    for p in PITCHES:
                        banks = {}
                        for row, sl in acc:
                            a = row * p + sl
                            banks.setdefault(a % 32, set()).add(a)
                        waves[p] += max(len(v) for v in banks.values())
                    b = rng.randint(32, size=len(acc))
                    waves["random"] += np.bincount(b, minlength=32).max()
        o.actor_step_philox(act, K, use_remi=True)
    print("requests (warp-level x loads of the car loop, first tick of 30 actor steps in steady state): %d" % req)
    for k, v in waves.items():
        print("pitch %-7s wavefronts per request %.3f   conflict share of wavefronts %.1f %%" % (k, v / req, 100.0 * (v - req) / v))


if __name__ == "__main__":
    main()
