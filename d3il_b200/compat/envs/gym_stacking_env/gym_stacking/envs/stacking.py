"""Reference import path of ``CubeStacking_Env`` (environments/d3il/envs/gym_stacking_env/gym_stacking/envs/stacking.py) on the batched CUDA backend."""
from d3il_b200.compat.gym_envs import CubeStacking_Env  # noqa: F401
