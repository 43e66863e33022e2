"""EnvPool: single-env proxies over one batched VecTrafficEnv, for learners written against one env.

The reference's learners each own an env and step it from their own loop - A3C from FLAGS.threads
Python threads (algorithms/a3c.py:66-72, 110-137; rollout contract `epoch`, :52-63), DQN / DRQN / PG from
one loop (qlearn.py:97-104, qrnn.py:108-118, polgrad_rnn.py:6-15).  `EnvPool(num_slots).slot(i)` gives
each of them an object with that single-env API (reset / step / action_space / observation_space /
reward_size / unwrapped.cars_on_roads / unwrapped.graph) while the device advances every slot in ONE
kernel launch per actor step: a step() call blocks until all slots that take part in the current round
have submitted their action, the last one launches the batch, everyone gets their row.

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

    def reset(self):
        self.pool._reset_slot(self.index)
        return self.step(self.action_space.sample())[0]

    def step(self, action):
        return self.pool._submit(self.index, action)

    def cars_on_roads(self):
        return self.pool._cars(self.index)

    def close(self):
        self.pool.leave(self.index)


class EnvPool(object):
    def __init__(self, num_slots, **vec_kwargs):
        vec_kwargs.setdefault("remi", True)
        self.vec = VecTrafficEnv(num_envs=num_slots, **vec_kwargs)
        self.num_slots = num_slots
        self._cv = threading.Condition()
        self._active = set(range(num_slots))      # slots whose owner is still stepping
        self._pending = {}                        # slot -> action of the current round
        self._round = 0
        self._results = {}
        self._actions = np.zeros((num_slots, self.vec.intersections), np.uint8)
        self._slots = [EnvSlot(self, i) for i in range(num_slots)]
        self.launches = 0

    def slot(self, i):
        return self._slots[i]

    # ---- called by the slots
    def _reset_slot(self, i):
        with self._cv:
            mask = np.zeros(self.num_slots, np.uint8)
            mask[i] = 1
            phases = np.random.randint(2, size=(self.num_slots, self.vec.intersections))
            self.vec.reset(mask=mask, init_phase=phases)

    def _cars(self, i):
        with self._cv:
            return self.vec.cars_on_roads()[i].copy()

    def leave(self, i):
        """The owner of slot i stops stepping (its env no longer gates the rounds)."""
        with self._cv:
            self._active.discard(i)
            self._pending.pop(i, None)
            self._maybe_launch()

    def _maybe_launch(self):
        if self._active and set(self._pending) >= self._active:
            for i, a in self._pending.items():
                self._actions[i] = np.asarray(a).astype(bool).reshape(-1)
            obs, rew, done = self.vec.step(self._actions)
            self.launches += 1
            self._results = {i: (obs[i].copy(), rew[i].copy(), bool(done[i]), None) for i in self._pending}
            self._pending = {}
            self._round += 1
            self._cv.notify_all()

    def _submit(self, i, action):
        with self._cv:
            if i not in self._active:
                self._active.add(i)
            my_round = self._round
            self._pending[i] = action
            self._maybe_launch()
            while self._round == my_round:
                self._cv.wait()
            return self._results[i]
