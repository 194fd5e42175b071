from .stacking import CubeStacking_Env  # noqa: F401
