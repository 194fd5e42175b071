"""Reference import path of ``Block_Push_Env`` (environments/d3il/envs/gym_pushing_env/gym_pushing/envs/pushing.py) on the batched CUDA backend."""
from d3il_b200.compat.gym_envs import Block_Push_Env  # noqa: F401
