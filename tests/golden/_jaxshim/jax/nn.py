import math
import numpy as _np
from . import numpy as jnp
from . import random as _random


def relu(x):
    return jnp._down(_np.maximum(x, 0))


def sigmoid(x):
    x = _np.asarray(x)
    return jnp._down((1.0 / (1.0 + _np.exp(-x))).astype(x.dtype))


def softplus(x):
    x = _np.asarray(x)
    return jnp._down(_np.logaddexp(x, _np.zeros((), x.dtype)).astype(x.dtype))


def tanh(x):
    return jnp._down(_np.tanh(x))


def softmax(x, axis=-1):
    e = _np.exp(x - _np.max(x, axis=axis, keepdims=True))
    return jnp._down(e / e.sum(axis=axis, keepdims=True))


class initializers:
    @staticmethod
    def glorot_uniform():
        def init(key, shape, dtype=_np.float32):
            a = math.sqrt(6.0 / (shape[0] + shape[1]))
            return jnp._down(_random._rs(key).uniform(-a, a, size=shape).astype(_np.float32))
        return init

    xavier_uniform = glorot_uniform

    @staticmethod
    def normal(stddev=1e-2):
        def init(key, shape, dtype=_np.float32):
            return jnp._down((_random._rs(key).normal(size=shape) * stddev).astype(_np.float32))
        return init

    @staticmethod
    def zeros(key, shape, dtype=_np.float32):
        return jnp._down(_np.zeros(shape, _np.float32))
