"""Reference import path of ``Sorting_Env`` (environments/d3il/envs/gym_sorting_env/gym_sorting/envs/sorting.py) on the batched CUDA backend."""
from d3il_b200.compat.gym_envs import Sorting_Env  # noqa: F401
