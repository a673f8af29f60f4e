import functools

import torch
from tensorflow.compat.v1 import _f
from tensorflow_probability import distributions as tfd
from .interceptor import interceptable


class RandomVariable(object):
    __array_ufunc__ = None      # numpy defers to our reflected operators (u * b1 with u an ndarray)
    __array_priority__ = 1000
    def __init__(self, distribution, sample_shape=(), value=None):
        self.distribution = distribution
        self._value = _f(value) if value is not None else distribution.sample()

    @property
    def value(self):
        return self._value

    @property
    def shape(self):
        return tuple(self._value.shape)

    def _arp_value(self):
        return self._value

    def __add__(self, o): return self._value + _f(o)
    def __radd__(self, o): return _f(o) + self._value
    def __sub__(self, o): return self._value - _f(o)
    def __rsub__(self, o): return _f(o) - self._value
    def __mul__(self, o): return self._value * _f(o)
    def __rmul__(self, o): return _f(o) * self._value
    def __truediv__(self, o): return self._value / _f(o)
    def __neg__(self): return -self._value
    def __getitem__(self, k): return self._value[k]


def _make_random_variable(distribution_cls):
    @interceptable
    @functools.wraps(distribution_cls, assigned=("__module__", "__name__"), updated=())
    def func(*args, **kwargs):
        sample_shape = kwargs.pop("sample_shape", ())
        value = kwargs.pop("value", None)
        return RandomVariable(distribution=distribution_cls(*args, **kwargs), sample_shape=sample_shape, value=value)
    return func


Normal = _make_random_variable(tfd.Normal)
Bernoulli = _make_random_variable(tfd.Bernoulli)
Gamma = _make_random_variable(tfd.Gamma)
TransformedDistribution = _make_random_variable(tfd.TransformedDistribution)
MultivariateNormalDiag = None
