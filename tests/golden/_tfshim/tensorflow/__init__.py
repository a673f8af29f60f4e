from .compat import v1  # noqa: F401
