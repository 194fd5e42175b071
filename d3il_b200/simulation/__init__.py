"""Drop-in rollout harnesses with the reference's class names and constructor signatures (``simulation/*_sim.py``)."""
from .aligning_sim import Aligning_Sim  # noqa: F401
from .avoiding_sim import Avoiding_Sim  # noqa: F401
from .base_sim import BaseSim  # noqa: F401
from .pushing_sim import Pushing_Sim  # noqa: F401
from .sorting_sim import Sorting_Sim  # noqa: F401
from .stacking_sim import Stacking_Sim  # noqa: F401
from .inserting_sim import Inserting_Sim  # noqa: F401
