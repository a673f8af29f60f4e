"""tfd.* stand-ins: TFP log_prob formulas in torch float64."""
import math

import torch
from tensorflow.compat.v1 import _f


class Distribution(object):
    def sample(self):
        raise NotImplementedError

    @property
    def parameters(self):
        return dict(self._params)


class Normal(Distribution):
    def __init__(self, loc=None, scale=None, name=None, **kw):
        self.loc, self.scale, self.name = _f(loc), _f(scale), name
        self._params = dict(loc=loc, scale=scale, name=name)

    def log_prob(self, x):
        u = (_f(x) - self.loc) / self.scale
        return -0.5 * u * u - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi)

    def sample(self):
        return torch.zeros(torch.broadcast_shapes(self.loc.shape, self.scale.shape), dtype=torch.float64)


class Bernoulli(Distribution):
    def __init__(self, logits=None, name=None, **kw):
        self.logits, self.name = _f(logits), name

    def log_prob(self, y):
        y, eta = _f(y), self.logits            # -sigmoid_cross_entropy_with_logits(labels=y, logits=eta)
        return y * eta - torch.clamp(eta, min=0.0) - torch.log1p(torch.exp(-torch.abs(eta)))

    def sample(self):
        return torch.zeros_like(self.logits)


class Gamma(Distribution):
    def __init__(self, concentration, rate, name=None, **kw):
        self.concentration, self.rate = _f(concentration), _f(rate)

    def log_prob(self, x):
        x = _f(x)
        a, b = self.concentration, self.rate
        return a * torch.log(b) - torch.lgamma(a) + (a - 1.0) * torch.log(x) - b * x

    def sample(self):
        return torch.ones(torch.broadcast_shapes(self.concentration.shape, self.rate.shape), dtype=torch.float64)


class TransformedDistribution(Distribution):
    def __init__(self, distribution, bijector=None, name=None, **kw):
        self.distribution, self.bijector, self.name = distribution, bijector, name

    def log_prob(self, y):
        y = _f(y)
        x = self.bijector.inverse(y)
        return self.distribution.log_prob(x) + self.bijector.inverse_log_det_jacobian(y)

    def sample(self):
        return self.bijector.forward(self.distribution.sample())
