"""tfb.* stand-ins."""
import torch
from tensorflow.compat.v1 import _f


class Exp(object):
    def forward(self, x):
        return torch.exp(_f(x))

    def inverse(self, y):
        return torch.log(_f(y))

    def inverse_log_det_jacobian(self, y):
        return -torch.log(_f(y))

    def forward_log_det_jacobian(self, x):
        return _f(x)


class Invert(object):
    def __init__(self, bijector):
        self.bijector = bijector

    def forward(self, x):
        return self.bijector.inverse(x)

    def inverse(self, y):
        return self.bijector.forward(y)

    def inverse_log_det_jacobian(self, y):
        return self.bijector.forward_log_det_jacobian(y)


class AffineScalar(object):
    def __init__(self, shift=None, scale=None):
        self.shift = 0.0 if shift is None else _f(shift)
        self.scale = 1.0 if scale is None else _f(scale)

    def forward(self, x):
        return self.shift + self.scale * _f(x)

    def inverse(self, y):
        return (_f(y) - self.shift) / self.scale
