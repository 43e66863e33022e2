"""GSpace: the multi-agent "generic space" every agent of the reference talks to
(reference gym_traffic/spaces/gspace.py:4-22).  Semantics preserved exactly, including that
sample() draws from the GLOBAL numpy RNG with the dtype of `limit`."""
import gym
import numpy as np


class GSpace(gym.Space):
    def __init__(self, shape, limit):
        self.shape = shape
        self.limit = limit
        self.size = int(np.prod(shape))

    def sample(self):
        return np.random.randint(self.limit, size=self.shape, dtype=self.limit.dtype)

    def contains(self, x):
        return x.shape == self.shape

    def empty(self):
        return np.empty(self.shape, dtype=self.limit.dtype)

    def to_action(self, a):
        return np.reshape(a, self.shape).astype(self.limit.dtype)

    def replicated(self, n):
        return GSpace([n] + self.shape, self.limit)

    def __repr__(self):
        return "GSpace(%r, %r)" % (self.shape, self.limit)
