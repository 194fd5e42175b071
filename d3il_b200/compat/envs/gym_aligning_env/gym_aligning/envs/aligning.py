"""Reference import path of ``Robot_Push_Env`` (environments/d3il/envs/gym_aligning_env/gym_aligning/envs/aligning.py) on the batched CUDA backend."""
from d3il_b200.compat.gym_envs import Robot_Push_Env  # noqa: F401
