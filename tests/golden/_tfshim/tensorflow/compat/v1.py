"""`import tensorflow.compat.v1 as tf` stand-in (torch float64)."""
import builtins
import types

import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
newaxis = None
DT = torch.float64


class ShimTensor(torch.Tensor):
    """torch.Tensor that, like a tf.Tensor, accepts numpy arrays / python lists as the other operand."""

    def __mul__(self, o): return torch.Tensor.__mul__(self, _fo(o))
    def __rmul__(self, o): return torch.Tensor.__mul__(self, _fo(o))
    def __add__(self, o): return torch.Tensor.__add__(self, _fo(o))
    def __radd__(self, o): return torch.Tensor.__add__(self, _fo(o))
    def __sub__(self, o): return torch.Tensor.__sub__(self, _fo(o))
    def __rsub__(self, o): return torch.Tensor.__sub__(_fo(o), self)
    def __truediv__(self, o): return torch.Tensor.__truediv__(self, _fo(o))
    def __rtruediv__(self, o): return torch.Tensor.__truediv__(_fo(o), self)


def _fo(o):
    """operand of a tensor operator: python scalars are exact (they are the stand-in's own
    constants, e.g. 0.5*log(2 pi)); arrays / lists are model data and enter as float32."""
    if isinstance(o, (int, float)):
        return torch.tensor(float(o), dtype=DT)
    return _f(o)


def _t(x):
    if hasattr(x, "_arp_value"):
        x = x._arp_value()
    if torch.is_tensor(x):
        return x if not x.dtype.is_floating_point else x.to(DT)
    a = np.array(x)
    if a.dtype.kind in "iub":
        return torch.as_tensor(a)
    # python / numpy constants enter a TF1 graph as float32 tensors: round once, then compute in float64
    return torch.as_tensor(a.astype(np.float32).astype(np.float64), dtype=DT)


def _f(x):
    return _t(x).to(DT).as_subclass(ShimTensor)


def convert_to_tensor(x, dtype=None):
    return _f(x)


constant = convert_to_tensor
identity = lambda x: _t(x)
ones = lambda shape, dtype=None: torch.ones(tuple(np.atleast_1d(shape).tolist()) if not isinstance(shape, int) else (shape,), dtype=DT)
zeros = lambda shape, dtype=None: torch.zeros(tuple(np.atleast_1d(shape).tolist()) if not isinstance(shape, int) else (shape,), dtype=DT)
ones_like = lambda x: torch.ones_like(_f(x))
zeros_like = lambda x: torch.zeros_like(_f(x))
exp = lambda x: torch.exp(_f(x))
log = lambda x: torch.log(_f(x))
sigmoid = lambda x: torch.sigmoid(_f(x))
pow = lambda x, y: torch.pow(_f(x), _f(y))
multiply = lambda x, y: _f(x) * _f(y)
negative = lambda x: -_f(x)
shape = lambda x=None, input=None: tuple(_t(x if x is not None else input).shape)
stack = lambda xs, axis=0: torch.stack([_f(x) for x in xs], dim=axis)
concat = lambda xs, axis: torch.cat([_f(x) for x in xs], dim=axis)
reshape = lambda x, s: _f(x).reshape(tuple(int(i) for i in s))
expand_dims = lambda x, axis: _f(x).unsqueeze(axis)
matmul = lambda a, b: _f(a) @ _f(b)
einsum = lambda eq, *xs: torch.einsum(eq, *[_f(x) for x in xs])


def reduce_sum(input_tensor=None, axis=None, **kw):
    x = _f(input_tensor)
    return x.sum() if axis is None else x.sum(dim=axis)


def one_hot(indices, depth, **kw):
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.int64)
    out = torch.zeros(tuple(idx.shape) + (int(depth),), dtype=DT)
    ok = (idx >= 0) & (idx < int(depth))          # tf.one_hot: out-of-range index -> all-zero row
    rows = torch.nonzero(ok, as_tuple=True)
    out[rows + (idx[ok],)] = 1.0
    return out


def get_variable(name=None, initializer=None, **kw):
    return _f(initializer)


def reset_default_graph():
    pass


def custom_gradient(f):
    return f


class _NN:
    softplus = staticmethod(lambda x: torch.nn.functional.softplus(_f(x)))


nn = _NN()


class _Random:
    @staticmethod
    def normal(shape, dtype=None, **kw):
        return torch.zeros(tuple(shape), dtype=DT)


random = _Random()


class _Linalg:
    class LinearOperator(object):
        pass


linalg = _Linalg()


class _GFile:
    GFile = staticmethod(lambda path, mode="r": builtins.open(path, mode))
    exists = staticmethod(lambda p: __import__("os").path.exists(p))


io = types.SimpleNamespace(gfile=_GFile)


class _FlagValues(object):
    def __init__(self):
        object.__setattr__(self, "_v", {})

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, "_v")[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        object.__getattribute__(self, "_v")[k] = v


class _Flags(object):
    FLAGS = _FlagValues()

    def _define(self, name, default=None, help=None, **kw):
        self.FLAGS._v.setdefault(name, default)

    DEFINE_string = DEFINE_boolean = DEFINE_integer = DEFINE_float = DEFINE_list = _define


app = types.SimpleNamespace(flags=_Flags())
