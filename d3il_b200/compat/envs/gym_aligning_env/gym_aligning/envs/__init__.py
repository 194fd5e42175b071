from .aligning import Robot_Push_Env  # noqa: F401
