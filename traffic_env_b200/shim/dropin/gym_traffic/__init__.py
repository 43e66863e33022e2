"""Drop-in gym_traffic package backed by the B200 simulator (traffic_env_b200).

Keeps what reference gym_traffic/__init__.py:6-23 sets up, because the agents and wrappers depend
on it: `gym.Env.step` renders first when `rendering` is set and then calls `_step` (so wrappers
that run several ticks per actor step can render the inner ticks), every env and wrapper carries a
`reward_size` (the traffic reward is a vector, one entry per intersection), and 'traffic-v0' is
registered.  The reference's own wrappers/ and algorithms/ packages keep working on top: if a
reference checkout is known (TRAFFIC_ENV_REFERENCE or on sys.path) its gym_traffic directory is
appended to this package's search path, so `gym_traffic.wrappers.*` / `gym_traffic.algorithms.*`
resolve there while `gym_traffic.envs` / `gym_traffic.spaces` resolve here.
"""
import os
import sys

import gym
from gym.envs.registration import register


def _render_then_step(self, action):
    if self.rendering:
        self.render()
    return self._step(action)


gym.Env.rendering = False
gym.Env.step = _render_then_step
gym.Env.reward_size = 1

if not getattr(gym.Wrapper.__init__, "_copies_reward_size", False):
    _wrapper_init = gym.Wrapper.__init__

    def _init_with_reward_size(self, env):
        _wrapper_init(self, env)
        self.reward_size = env.reward_size

    _init_with_reward_size._copies_reward_size = True
    gym.Wrapper.__init__ = _init_with_reward_size

register(id="traffic-v0", entry_point="gym_traffic.envs:TrafficEnv")

_here = os.path.dirname(os.path.abspath(__file__))
_candidates = [os.environ.get("TRAFFIC_ENV_REFERENCE")] + list(sys.path)
for _p in _candidates:
    if not _p:
        continue
    _d = os.path.join(_p, "gym_traffic")
    if os.path.isdir(os.path.join(_d, "algorithms")) and os.path.abspath(_d) != _here and _d not in __path__:
        __path__.append(_d)
        break
