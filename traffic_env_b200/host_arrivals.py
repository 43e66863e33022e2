"""Host-side restatement of the reference's MT19937 arrival generators, for the single-env
drop-in (gym_traffic.envs.TrafficEnv): with the same seed it feeds the device the very schedule
the reference would have produced, so trajectories are reproducible against the reference.

Reference: gym_traffic/envs/traffic_env.py:160-164 (poisson), :167-176 (regular), :274-283
(add_new_cars: the entry road of every car is `rand.choice(entrypoints)`, drawn from the SAME
RandomState right after the generator yields the car).  The order in which the RandomState is
consumed - exponential, randint(1), choice, exponential, ... - is what has to be preserved.
"""
import math

import numpy as np


class ArrivalStream(object):
    """Per-tick lists of entry roads.  `tick_roads()` returns the ordered roads of the next tick."""

    def __init__(self, seed, entrypoints, cars_per_sec, rate, poisson=True):
        self.rand = np.random.RandomState(seed)
        self.entrypoints = np.asarray(entrypoints)
        self.cars_per_sec = float(cars_per_sec)
        self.rate = float(rate)
        self._gen = self._poisson() if poisson else self._regular()

    def _poisson(self):
        scale = 1 / (self.cars_per_sec * self.rate)
        while True:
            for _ in range(round(self.rand.exponential(scale))):
                yield False            # an empty tick boundary
            self.rand.randint(1)       # the reference picks an archetype here (there is one); keeps the stream aligned
            yield True                 # a car

    def _regular(self):
        per_tick = self.cars_per_sec * self.rate
        every = round(1 / per_tick)
        burst = math.ceil(per_tick)
        i = 0
        while True:
            if every == 0 or i % every == 0:
                for _ in range(burst):
                    yield True
            yield False
            i += 1

    def tick_roads(self):
        roads = []
        while next(self._gen):
            roads.append(int(self.rand.choice(self.entrypoints)))
        return roads

    def window(self, ticks):
        return [self.tick_roads() for _ in range(ticks)]
