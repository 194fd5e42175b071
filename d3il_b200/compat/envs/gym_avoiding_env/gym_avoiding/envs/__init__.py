from .avoiding import ObstacleAvoidanceEnv  # noqa: F401
