from .pushing import Block_Push_Env  # noqa: F401
