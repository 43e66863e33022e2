"""Manual long soak (not collected by pytest): the long-horizon parity test of test_gpu_philox_sharding.py at sizes
that take the oracle tens of seconds (multi-step launches with the controller in the kernel - one or two greedy decisions per launch - and single steps) - 24 envs x 6000 ticks of the 10x10 bench workload, 256 envs x 15000 ticks of
the default grid, 7x7 / 13x13 grids, heavy arrivals.  Every actor step and the final ring state bit for bit.

  python tests/big_soak_run.py          (on a B200; ~40 s)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_philox_sharding import test_long_horizon_soak_vs_oracle as soak  # noqa: E402

if __name__ == "__main__":
    for args in [(10, 10, 500.0, 24, 600, 0.12, 6), (3, 3, 250.0, 256, 1500, 0.12, 6), (7, 7, 300.0, 16, 600, 0.15, True),
                 (13, 13, 200.0, 4, 300, 0.1, False), (3, 3, 250.0, 64, 600, 0.3, True), (10, 10, 500.0, 8, 600, 0.12, False),
                 (3, 3, 250.0, 128, 900, 0.12, False)]:
        t0 = time.time()
        soak(*args)
        print("ok", args, "%.1f s" % (time.time() - t0), flush=True)
