from .running_norm import RunningNorm  # noqa: F401
