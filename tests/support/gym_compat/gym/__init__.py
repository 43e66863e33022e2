"""Minimal stand-in for the 2016/2017-era ``gym`` API (0.7 - 0.9) that
``gym_traffic`` was written against.

The real ``gym`` package is not installed in this image and modern
``gymnasium`` dropped the ``_step``/``_reset`` indirection the reference relies
on (reference: gym_traffic/__init__.py:6-18 monkey-patches ``gym.Env.step`` and
``gym.Wrapper.__init__``; traffic_test.py:27-76 subclasses ``gym.Wrapper`` and
``gym.RewardWrapper``).  Only the surface the reference touches is provided:

* ``Env``: ``step -> _step``, ``reset -> _reset``, ``render -> _render``,
  ``close``, ``seed``, ``unwrapped``, ``action_space``, ``observation_space``.
* ``Wrapper``/``RewardWrapper``/``ObservationWrapper``/``ActionWrapper``.
* ``Space``, ``spaces.Discrete``, ``spaces.Box``.
* ``make`` / ``envs.registration.register`` with ``module:Class`` entry points.

It is host-side plumbing only; nothing in here is on the simulation path.
"""
import importlib

__version__ = "0.9.compat"


class Space(object):
    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError


class Env(object):
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    # old-style indirection: public method -> underscore method
    def step(self, action):
        return self._step(action)

    def reset(self):
        return self._reset()

    def render(self, mode="human", close=False):
        return self._render(mode=mode, close=close)

    def close(self):
        return self._close()

    def seed(self, seed=None):
        return self._seed(seed)

    def _step(self, action):
        raise NotImplementedError

    def _reset(self):
        raise NotImplementedError

    def _render(self, mode="human", close=False):
        return None

    def _close(self):
        return None

    def _seed(self, seed=None):
        return []

    @property
    def unwrapped(self):
        return self

    def __str__(self):
        return "<%s instance>" % type(self).__name__


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.reward_range = env.reward_range
        self.metadata = env.metadata

    def _step(self, action):
        return self.env.step(action)

    def _reset(self):
        return self.env.reset()

    def _render(self, mode="human", close=False):
        return self.env.render(mode=mode, close=close)

    def _close(self):
        return self.env.close()

    def _seed(self, seed=None):
        return self.env.seed(seed)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def spec(self):
        return getattr(self.env, "spec", None)


class ObservationWrapper(Wrapper):
    def _reset(self):
        return self._observation(self.env.reset())

    def _step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self._observation(obs), reward, done, info

    def observation(self, observation):
        return self._observation(observation)

    def _observation(self, observation):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def _step(self, action):
        obs, reward, done, info = self.env.step(action)
        return obs, self._reward(reward), done, info

    def reward(self, reward):
        return self._reward(reward)

    def _reward(self, reward):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def _step(self, action):
        return self.env.step(self._action(action))

    def action(self, action):
        return self._action(action)

    def _action(self, action):
        raise NotImplementedError


from gym import spaces  # noqa: E402
from gym.envs import registration  # noqa: E402
from gym.envs.registration import make, register, spec  # noqa: E402

__all__ = ["Env", "Space", "Wrapper", "ObservationWrapper", "RewardWrapper",
           "ActionWrapper", "make", "register", "spec", "spaces"]
