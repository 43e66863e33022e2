from gym.envs import registration  # noqa: F401
from gym.envs.registration import make, register, spec, registry  # noqa: F401
