"""Host-side restatement of the reference's MT19937 arrival generators, for the single-env
drop-in (gym_traffic.envs.TrafficEnv): with the same seed it feeds the device the very schedule
the reference would have produced, so trajectories are reproducible against the reference.

Reference: gym_traffic/envs/traffic_env.py:160-164 (poisson), :167-176 (regular), :274-283
(add_new_cars: the entry road of every car is `rand.choice(entrypoints)`, drawn from the SAME
RandomState right after the generator yields the car).  The order in which the RandomState is
consumed - exponential, randint(1), choice, exponential, ... - is what has to be preserved.

The generators are written as an explicit state machine (not Python generators) so the stream can be
snapshotted and rewound: the drop-in hands the device a window of future ticks at a time, while the
reference draws lazily - when the entry set changes (reset_entrypoints) the not-yet-executed part of
the window is re-drawn from the rewound stream with the new entry set.
"""
import copy
import math

import numpy as np


class ArrivalStream(object):
    """Per-tick lists of entry roads.  `tick_roads()` returns the ordered roads of the next tick."""

    def __init__(self, seed, entrypoints, cars_per_sec, rate, poisson=True):
        self.rand = np.random.RandomState(seed)
        self.entrypoints = np.asarray(entrypoints)
        self.poisson = bool(poisson)
        per_tick = float(cars_per_sec) * float(rate)
        self.scale = 1 / per_tick                    # poisson(): lam, fixed when the generator starts
        self.every = round(1 / per_tick)             # regular(): ticks_per_car
        self.burst = math.ceil(per_tick)             # regular(): cars_per_tick_int
        self._skip = None                            # poisson: empty ticks left before the next car (None: not drawn yet)
        self._i = 0                                  # regular: tick counter

    def tick_roads(self):
        roads = []
        if self.poisson:
            while True:
                if self._skip is None:
                    self._skip = round(self.rand.exponential(self.scale))
                if self._skip > 0:
                    self._skip -= 1                  # one `None` of the generator: the tick ends
                    return roads
                self.rand.randint(1)                 # the reference picks an archetype here (there is one)
                roads.append(int(self.rand.choice(self.entrypoints)))
                self._skip = None
        if self.every == 0 or self._i % self.every == 0:
            for _ in range(self.burst):
                roads.append(int(self.rand.choice(self.entrypoints)))
        self._i += 1
        return roads

    def window(self, ticks):
        return [self.tick_roads() for _ in range(ticks)]

    def snapshot(self):
        return (self.rand.get_state(), self._skip, self._i, self.entrypoints.copy())

    def restore(self, snap):
        state, self._skip, self._i, entry = snap
        self.rand.set_state(copy.deepcopy(state))
        self.entrypoints = entry.copy()
