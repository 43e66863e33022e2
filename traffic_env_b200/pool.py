"""EnvPool: single-env proxies over one batched VecTrafficEnv, for learners written against one env.

The reference's learners each own an env and step it from their own loop - A3C from FLAGS.threads
Python threads (algorithms/a3c.py:66-72, 110-137; rollout contract `epoch`, :52-63), DQN / DRQN / PG from
one loop (qlearn.py:97-104, qrnn.py:108-118, polgrad_rnn.py:6-15).  `EnvPool(num_slots).slot(i)` gives
each of them an object with that single-env API (reset / step / action_space / observation_space /
reward_size / unwrapped.cars_on_roads / unwrapped.graph) while the device advances the slots in batches.

No slot waits for another slot's owner: a step() call queues its action; whichever caller finds no launch in
flight becomes the leader, takes EVERY action queued so far (at least its own) and advances exactly those slots
with one te_step_masked launch (the other slots' state and arrival streams are untouched); callers that arrive
while a launch is in flight are served by the next one.  A slow learner thread therefore only delays itself, and
the batch size adapts to how many threads are ready.  The queue, the leader election and the launch live in the
library (te_pool_step, csrc/te_pool.cpp) and run without the GIL.

Semantics per slot are those of Remi(Repeater(K)) on the reference env (traffic_test.py:27-64): reset() is
TrafficEnv._reset followed by one step with a random action whose observation is returned (:34-36).
"""
import numpy as np

from .vec_env import VecTrafficEnv


def _gspace_class():
    """The drop-in's GSpace (shim/dropin/gym_traffic/spaces/gspace.py), loaded by path so that the pool does not
    need the `gym_traffic` package (whose __init__ imports gym) to be installed."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim", "dropin", "gym_traffic", "spaces", "gspace.py")
    spec = importlib.util.spec_from_file_location("traffic_env_b200._gspace", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.GSpace


class _Graph(object):
    def __init__(self, vec):
        self.m, self.n = vec.m, vec.n
        self.train_roads, self.roads, self.intersections = vec.train_roads, vec.roads, vec.intersections
        self.nexts, self.dest, self.phases, self.entrypoints = vec.nexts, vec.dest, vec.phases, vec.entrypoints


class EnvSlot(object):
    """What one learner thread sees: the gym-style API of a single wrapped traffic env."""

    def __init__(self, pool, index):
        GSpaceLike = _gspace_class()
        self.pool, self.index = pool, index
        v = pool.vec
        self.action_space = GSpaceLike([v.intersections], np.int32(2))
        self.observation_space = GSpaceLike([v.obs_len], np.float32(1))
        self.reward_size = v.intersections
        self.graph = _Graph(v)
        self.rendering = False

    @property
    def unwrapped(self):
        return self

    def reset(self, init_phase=None, first_action=None):
        """Repeater._reset (traffic_test.py:34-36): TrafficEnv._reset, then one step with a random action whose
        observation is returned.  init_phase / first_action (extensions) replace the two random draws."""
        self.pool._reset_slot(self.index, init_phase)
        return self.step(self.action_space.sample() if first_action is None else first_action)[0]

    def step(self, action):
        return self.pool._submit(self.index, action)

    def cars_on_roads(self):
        return self.pool._cars(self.index)

    def close(self):
        self.pool.leave(self.index)


class EnvPool(object):
    """num_slots env instances in one batched handle, stepped through the library's slot pool (te_pool_*,
    csrc/te_pool.cpp): queueing, leader election and the masked launch all run outside the Python GIL."""

    def __init__(self, num_slots, linger=150e-6, **vec_kwargs):
        """linger: seconds a would-be leader waits ONCE for more actions when fewer slots are queued than the previous
        launch served (bigger batches when all learners are fast; a slow learner costs the others at most this much
        per step, never a whole round)."""
        import ctypes as C
        from ._lib import check
        vec_kwargs.setdefault("remi", True)
        self.vec = VecTrafficEnv(num_envs=num_slots, **vec_kwargs)
        self.num_slots = num_slots
        self._C, self._L = C, self.vec._L
        self._p = C.c_void_p()
        check(self._L.te_pool_create(self.vec._h, self.vec.ticks_per_step, int(round(linger * 1e6)), C.byref(self._p)))
        self._slots = [EnvSlot(self, i) for i in range(num_slots)]

    def close(self):
        if getattr(self, "_p", None) is not None and self._p.value:
            self._L.te_pool_destroy(self._p)
            self._p = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def slot(self, i):
        return self._slots[i]

    def _fail(self):
        from ._lib import TrafficB200Error
        raise TrafficB200Error(self._L.te_pool_last_error(self._p).decode("utf-8", "replace"))

    @property
    def launches(self):
        return self._counters()[0]

    @property
    def stepped(self):
        """env actor steps served (sum of the batch sizes)"""
        return self._counters()[1]

    def _counters(self):
        a, b = self._C.c_uint64(), self._C.c_uint64()
        self._L.te_pool_counters(self._p, self._C.byref(a), self._C.byref(b))
        return a.value, b.value

    # ---- called by the slots
    def _reset_slot(self, i, init_phase=None):
        I = self.vec.intersections
        ph = np.random.randint(2, size=I) if init_phase is None else np.asarray(init_phase).astype(bool)
        ph = np.ascontiguousarray(ph, dtype=np.uint8)
        if self._L.te_pool_reset(self._p, int(i), ph.ctypes.data) < 0:
            self._fail()

    def _cars(self, i):
        out = np.empty(self.vec.roads, np.int32)
        if self._L.te_pool_cars(self._p, int(i), out.ctypes.data) < 0:
            self._fail()
        c = out[:self.vec.train_roads]
        return np.transpose(c.reshape(4, self.vec.m, self.vec.n), (1, 2, 0))

    def leave(self, i):
        """The owner of slot i stops stepping (kept for API compatibility: nobody waits for it anyway)."""

    def _submit(self, i, action):
        v = self.vec
        a = np.ascontiguousarray(np.asarray(action).astype(bool).reshape(-1), dtype=np.uint8)
        obs = np.empty(v.obs_len, np.float32)
        rew = np.empty(v.intersections, np.float32)
        done = self._C.c_uint8(0)
        if self._L.te_pool_step(self._p, int(i), a.ctypes.data, obs.ctypes.data, rew.ctypes.data, self._C.byref(done)) < 0:
            self._fail()
        return obs, rew, bool(done.value), None
