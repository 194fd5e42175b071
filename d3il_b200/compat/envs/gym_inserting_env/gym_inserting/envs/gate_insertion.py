"""Reference import path of ``Gate_Insertion_Env`` (environments/d3il/envs/gym_inserting_env/gym_inserting/envs/gate_insertion.py) on the batched CUDA backend."""
from d3il_b200.compat.gym_envs import Gate_Insertion_Env  # noqa: F401
