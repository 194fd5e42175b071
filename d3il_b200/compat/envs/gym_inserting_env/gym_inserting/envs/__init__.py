from .gate_insertion import Gate_Insertion_Env  # noqa: F401
