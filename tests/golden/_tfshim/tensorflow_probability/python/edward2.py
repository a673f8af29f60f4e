from tensorflow_probability.python.experimental.edward2 import *  # noqa: F401,F403
from tensorflow_probability.python.experimental.edward2 import (Bernoulli, Gamma, Normal, RandomVariable,  # noqa: F401
                                                                 TransformedDistribution, interceptable, interception, tape)
