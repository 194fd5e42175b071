from .sorting import Sorting_Env  # noqa: F401
