"""Reference import path of ``ObstacleAvoidanceEnv`` (environments/d3il/envs/gym_avoiding_env/gym_avoiding/envs/avoiding.py) on the batched CUDA backend."""
from d3il_b200.compat.gym_envs import ObstacleAvoidanceEnv  # noqa: F401
