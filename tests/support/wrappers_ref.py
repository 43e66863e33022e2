"""TEST SUPPORT - restatements of the reference launcher's host-side wrapper shims and env factory
(traffic_test.py:66-91, gym_traffic/wrappers/{warmup,history,gspace}.py), needed only because the reference checkout
is absent on the GPU box.  With the checkout present the reference's own files are used on top of the drop-in
(tests/test_reference_live.py).  The product package keeps only the two tick-loop wrappers it fuses
(traffic_env_b200/wrappers.py: Repeater, Remi).
"""
import gym
import numpy as np
from args import FLAGS

from traffic_env_b200.wrappers import Remi, Repeater


def WarmupWrapper(ignore_count):
    """`ignore_count` random-action steps after every reset (gym_traffic/wrappers/warmup.py:3-14)."""
    class WarmupWrapper(gym.Wrapper):
        def _reset(self):
            obs = self.env.reset()
            for _ in range(ignore_count):
                obs, _, done, _ = self.env.step(self.env.action_space.sample())
                assert not done, "Episode completed during warmup"
            return obs
    return WarmupWrapper


def HistoryWrapper(history_count):
    """Observation = the last `history_count` observations, oldest first (gym_traffic/wrappers/history.py:5-26)."""
    from collections import deque

    class HistoryWrapper(gym.Wrapper):
        def __init__(self, env):
            super(HistoryWrapper, self).__init__(env)
            self.history = deque(maxlen=history_count)
            self.observation_space = env.observation_space.replicated(history_count)

        def _reset(self):
            self.history.clear()
            self.history.append(self.env.reset())
            while len(self.history) < history_count:
                self.history.append(self.env.step(self.env.action_space.sample())[0])
            return np.stack(self.history)

        def _step(self, action):
            obs, reward, done, info = self.env.step(action)
            self.history.append(obs)       # maxlen drops the oldest
            return np.array(self.history), reward, done, info
    return HistoryWrapper


class LocalizeWrapper(gym.RewardWrapper):
    """Each intersection's reward = weighted mean of all rewards with its own counted `local_weight` times
    (traffic_test.py:66-69)."""
    def _reward(self, a):
        w = FLAGS.local_weight
        return np.mean(np.diag(a) * (w - 1) + a, axis=1) / w


class SquishReward(gym.RewardWrapper):
    """Scalar reward = mean over intersections (traffic_test.py:71-76)."""
    def __init__(self, env):
        super(SquishReward, self).__init__(env)
        self.reward_size = 1

    def _reward(self, a):
        return np.mean(a)


class UnGSpaceWrapper(gym.Wrapper):
    """Single-agent view: one Discrete action index unravelled into the per-intersection action tuple, mean reward
    (gym_traffic/wrappers/gspace.py:23-34)."""
    def __init__(self, env):
        super(UnGSpaceWrapper, self).__init__(env)
        from gym.spaces import Box, Discrete
        self.action_gspace = self.action_space
        self.observation_gspace = self.observation_space
        self.action_space = Discrete(self.action_gspace.size)
        self.observation_space = Box(0, self.action_gspace.limit, shape=self.observation_gspace.shape)

    def _step(self, action):
        obs, reward, done, info = self.env.step(np.unravel_index(action, self.action_gspace.shape))
        return obs, np.mean(reward), done, info


def make_env(m=3, n=3, length=250, seed=None, light_iterations=None, remi=None, warmup_lights=None, local_weight=None,
             squish_rewards=None, history=None, single_agent=None):
    """The reference's env factory (traffic_test.py:78-91) on the B200 env: Repeater -> [Warmup] -> [Remi] ->
    [Localize] -> [Squish] -> [History] -> [UnGSpace], each layer switched by the same flag as in the reference
    (arguments override FLAGS)."""
    import gym_traffic  # noqa: F401  registers traffic-v0
    from gym_traffic.envs.roadgraph import GridRoad

    def flag(value, name, default):
        if value is not None:
            return value
        try:
            return getattr(FLAGS, name)
        except AttributeError:
            return default

    env = gym.make('traffic-v0')
    env.set_graph(GridRoad(m, n, length))
    env.seed_generator(seed)
    env.reset_entrypoints()
    if light_iterations is None:
        light_iterations = flag(None, "light_iterations", None) or int(flag(None, "light_secs", 5) / FLAGS.rate)
    env = Repeater(light_iterations)(env)
    if flag(warmup_lights, "warmup_lights", 0) > 0:
        env = WarmupWrapper(flag(warmup_lights, "warmup_lights", 0))(env)
    if flag(remi, "remi", True):
        env = Remi(env)
    if flag(local_weight, "local_weight", 1) > 1:
        FLAGS.local_weight = flag(local_weight, "local_weight", 1)
        env = LocalizeWrapper(env)
    if flag(squish_rewards, "squish_rewards", False):
        env = SquishReward(env)
    if flag(history, "history", 1) > 1:
        env = HistoryWrapper(flag(history, "history", 1))(env)
    if flag(single_agent, "single_agent", False):
        env = UnGSpaceWrapper(env)
    return env
