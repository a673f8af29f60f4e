from .generated_random_variables import *  # noqa: F401,F403
from .generated_random_variables import Bernoulli, Gamma, Normal, RandomVariable, TransformedDistribution  # noqa: F401
from .interceptor import get_next_interceptor, interceptable, interception, tape  # noqa: F401
