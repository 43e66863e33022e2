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
the batch size adapts to how many threads are ready.

Semantics per slot are those of Remi(Repeater(K)) on the reference env (traffic_test.py:27-64): reset() is
TrafficEnv._reset followed by one step with a random action whose observation is returned (:34-36).
"""
import threading

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
    def __init__(self, num_slots, linger=150e-6, **vec_kwargs):
        """linger: seconds a would-be leader waits ONCE for more actions when fewer slots are queued than the previous
        launch served (bigger batches when all learners are fast; a slow learner costs the others at most this much
        per step, never a whole round)."""
        vec_kwargs.setdefault("remi", True)
        self._linger = float(linger)
        self._last_batch = 1
        self.vec = VecTrafficEnv(num_envs=num_slots, **vec_kwargs)
        self.num_slots = num_slots
        self._cv = threading.Condition()
        self._pending = {}                        # slot -> action, queued for the next launch
        self._results = {}                        # slot -> (obs, reward, done, info) of its last step, until fetched
        self._launching = False
        self._actions = np.zeros((num_slots, self.vec.intersections), np.uint8)
        self._mask = np.zeros(num_slots, np.uint8)
        self._slots = [EnvSlot(self, i) for i in range(num_slots)]
        self.launches = 0
        self.stepped = 0                          # env actor steps served (sum of batch sizes)

    def slot(self, i):
        return self._slots[i]

    # ---- called by the slots
    def _reset_slot(self, i, init_phase=None):
        with self._cv:
            while self._launching:                # the device state of slot i must not change under a launch
                self._cv.wait()
            mask = np.zeros(self.num_slots, np.uint8)
            mask[i] = 1
            phases = np.zeros((self.num_slots, self.vec.intersections), np.uint8)
            phases[i] = np.random.randint(2, size=self.vec.intersections) if init_phase is None else np.asarray(init_phase).astype(bool)
            self.vec.reset(mask=mask, init_phase=phases)

    def _cars(self, i):
        with self._cv:
            while self._launching:
                self._cv.wait()
            return self.vec.cars_on_roads()[i].copy()

    def leave(self, i):
        """The owner of slot i stops stepping (kept for API compatibility: nobody waits for it anyway)."""
        with self._cv:
            self._pending.pop(i, None)

    def _submit(self, i, action):
        with self._cv:
            self._pending[i] = np.asarray(action).astype(bool).reshape(-1)
            lingered = False
            while True:
                if i in self._results:
                    res = self._results.pop(i)
                    if isinstance(res, BaseException):
                        raise res
                    return res
                if not self._launching and i in self._pending:
                    if not lingered and self._linger > 0 and len(self._pending) < self._last_batch:
                        lingered = True
                        self._cv.wait(self._linger)       # others may queue (or lead) meanwhile
                        continue
                    batch, self._pending = self._pending, {}
                    self._launching = True
                    self._last_batch = len(batch)
                    break
                self._cv.wait()
        # leader: one masked launch for everything that was queued (the lock is released: others keep queueing)
        out = None
        try:
            self._mask[:] = 0
            for j, a in batch.items():
                self._actions[j] = a
                self._mask[j] = 1
            obs, rew, done = self.vec.step_masked(self._actions, self._mask)
            out = {j: (obs[j].copy(), rew[j].copy(), bool(done[j]), None) for j in batch}
        except BaseException as ex:               # every caller of the failed batch gets the error, nobody hangs
            out = {j: ex for j in batch}
        with self._cv:
            self._launching = False
            self.launches += 1
            self.stepped += len(batch)
            self._results.update(out)
            self._cv.notify_all()
            mine = self._results.pop(i)
        if isinstance(mine, BaseException):
            raise mine
        return mine
