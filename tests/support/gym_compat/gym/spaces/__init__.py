"""Discrete / Box: only what reference gym_traffic/wrappers/gspace.py:3,27-29
touches (``Discrete(n)``, ``Box(low, high, shape=...)``, ``.n``, ``.high``)."""
import numpy as np


class _Space(object):
    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError


class Discrete(_Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Box(_Space):
    def __init__(self, low, high, shape=None):
        if shape is None:
            self.low = np.asarray(low, dtype=np.float32)
            self.high = np.asarray(high, dtype=np.float32)
        else:
            self.low = np.full(shape, low, dtype=np.float32)
            self.high = np.full(shape, high, dtype=np.float32)
        self.shape = self.low.shape

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(np.float32)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def __repr__(self):
        return "Box%s" % (self.shape,)
