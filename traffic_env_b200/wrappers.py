"""Tick-loop wrappers of the reference launcher (traffic_test.py:27-64) for the drop-in TrafficEnv.

`Repeater(k)` runs k physics ticks per actor step and accumulates the observation; `Remi` replaces the
reward by TrafficEnv.remi_reward().  When Repeater sits directly on the B200 TrafficEnv and nothing
asks for per-tick rendering, its k ticks are ONE kernel launch (TrafficEnv.step_repeated); otherwise it
falls back to the reference's tick-by-tick loop over env.step, which is still the device path, one
launch per tick.  The remaining layers of the reference's factory (Warmup, Localize, Squish, History,
UnGSpace: thin host-side numpy shims in traffic_test.py / gym_traffic/wrappers) are out of scope: the
reference's own files stack on top of these two unchanged.
"""
import gym
import numpy as np
from args import FLAGS

from gym_traffic.spaces.gspace import GSpace


def _validate_mode():
    try:
        return FLAGS.mode == 'validate'
    except AttributeError:
        return False


def Repeater(repeat_count):
    class Repeater(gym.Wrapper):
        def __init__(self, env):
            super(Repeater, self).__init__(env)
            g = self.unwrapped.graph
            self.r, self.i = g.train_roads, g.intersections
            self.observation_space = GSpace([2 * self.r + self.i], np.float32(1))

        def _reset(self):
            super(Repeater, self)._reset()
            return self._step(self.action_space.sample())[0]

        def _light_times(self, action):
            # seconds each light that switches now has held its phase (validate-mode info, traffic_test.py:41-46)
            change = np.logical_xor(self.env.current_phase, action).astype(np.int32)
            held = ((self.env.elapsed + 1) * change).astype(np.float32) / 2
            return {'light_times': held[np.nonzero(held)]}

        def _step(self, action):
            info = self._light_times(action) if _validate_mode() else None
            base = self.env
            if hasattr(base, "step_repeated") and not base.rendering:
                obs, reward, done = base.step_repeated(action, repeat_count)
                return obs, reward.copy(), done, info
            total_obs = np.zeros(self.observation_space.shape, dtype=np.float32)
            total_reward, done = 0, False
            for _ in range(repeat_count):
                obs, reward, done, _ = base.step(action)
                total_obs[:self.r] += obs[:self.r]
                total_obs[self.r:2 * self.r] = obs[self.r:2 * self.r]
                total_obs[-self.i:] = obs[-self.i:] / 100 * (2 * obs[-2 * self.i:-self.i] - 1)
                total_reward += reward
                if done:
                    break
            return total_obs, total_reward, done, info
    return Repeater


class Remi(gym.Wrapper):
    def _step(self, action):
        obs, _, done, info = self.env.step(action)
        reward = self.unwrapped.remi_reward()
        return obs, reward, done, info
