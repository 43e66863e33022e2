"""Host-side definition of the counter-based arrival process.

The reference draws arrivals from Python generators over MT19937
(gym_traffic/envs/traffic_env.py:160-164, 274-283): k = round(Exp(scale = 1 /
(cars_per_sec * rate))) empty ticks, then one car on a uniformly chosen entry road,
repeat (k = 0 puts several cars in one tick).  The device keeps that process but
draws from Philox4x32-10 keyed by (seed, global env id), counter = draw index:
word 0 picks the gap by inverse CDF against the 32-bit thresholds below, word 1
picks the entry road as (word * n_entry) >> 32.  te_api.cu builds the same table
(build_gap_cdf); this Python copy exists so tests and the CPU oracle can share it.
"""
import math

import numpy as np


def gap_cdf(cars_per_tick):
    """T[k] = floor(2^32 * P(round(Exp) <= k)) = floor(2^32 * (1 - exp(-(k + 0.5) * cars_per_tick)))."""
    out = []
    if not cars_per_tick > 0:
        return np.zeros(0, np.uint32)
    for k in range(8192):
        cdf = -math.expm1(-(k + 0.5) * cars_per_tick)
        scaled = math.floor(cdf * 4294967296.0)
        if scaled >= 4294967295.0:
            break
        out.append(scaled)
    return np.asarray(out, dtype=np.uint32)
