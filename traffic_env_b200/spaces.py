"""GSpace without a gym dependency (same semantics as reference gym_traffic/spaces/gspace.py:4-22), for
code paths that must not require `gym` to be importable (EnvPool)."""
import numpy as np


class GSpaceLike(object):
    def __init__(self, shape, limit):
        self.shape, self.limit, self.size = shape, limit, int(np.prod(shape))

    def sample(self):
        return np.random.randint(self.limit, size=self.shape, dtype=self.limit.dtype)

    def contains(self, x):
        return x.shape == self.shape

    def empty(self):
        return np.empty(self.shape, dtype=self.limit.dtype)

    def to_action(self, a):
        return np.reshape(a, self.shape).astype(self.limit.dtype)

    def replicated(self, n):
        return GSpaceLike([n] + self.shape, self.limit)
