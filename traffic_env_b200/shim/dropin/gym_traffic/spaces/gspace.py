"""GSpace: the multi-agent "generic space" every agent of the reference talks to
(reference gym_traffic/spaces/gspace.py:4-22): a box of `shape` whose elements are integers (or floats) below
`limit`, typed by `limit`'s dtype.  Semantics preserved exactly - including that sample() draws from the GLOBAL
numpy RNG with that dtype (action sequences of seeded runs depend on it) and that `contains` compares the array
shape with the (list) shape as given."""
import numpy as np

try:
    from gym import Space as _Space
except ImportError:      # EnvPool (traffic_env_b200/pool.py) loads this file by path and works without gym
    _Space = object


class GSpace(_Space):
    def __init__(self, shape, limit):
        self.shape, self.limit = shape, limit
        self.size = int(np.prod(shape))

    @property
    def dtype(self):
        return self.limit.dtype

    def sample(self):
        draw = np.random.randint                       # module-level RandomState, like the reference (:14)
        return draw(self.limit, size=self.shape, dtype=self.dtype)

    def contains(self, x):
        return x.shape == self.shape

    def empty(self):
        return np.empty(self.shape, self.dtype)

    def to_action(self, a):
        return np.asarray(a).reshape(self.shape).astype(self.dtype)

    def replicated(self, n):
        return GSpace([n, *self.shape], self.limit)

    def __repr__(self):
        return "GSpace(%r, %r)" % (self.shape, self.limit)
