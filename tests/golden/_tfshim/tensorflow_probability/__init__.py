from . import bijectors, distributions  # noqa: F401
from .python.experimental import edward2  # noqa: F401
