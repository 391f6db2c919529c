"""Deterministic stand-ins for jax.random (threefry streams cannot be reproduced without JAX); every draw is
recorded in DRAWS so that the golden script can hand the same numbers to the oracle."""
import numpy as _np
from . import numpy as jnp

DRAWS = []


def PRNGKey(seed):
    return _np.array([0, int(seed) & 0x7FFFFFFF], dtype=_np.uint32)


def _rs(key):
    return _np.random.RandomState(int(_np.asarray(key).astype(_np.uint64).sum()) & 0x7FFFFFFF)


def split(key, num=2):
    base = int(_np.asarray(key).astype(_np.uint64).sum())
    return [_np.array([i + 1, (base * 1000003 + 7919 * (i + 1)) & 0x7FFFFFFF], dtype=_np.uint32) for i in range(num)]


def randint(key, shape, minval, maxval):
    out = jnp._down(_rs(key).randint(minval, maxval, size=tuple(shape)))
    DRAWS.append(("randint", out))
    return out


def uniform(key, shape, dtype=None, minval=0.0, maxval=1.0):
    out = jnp._down((_rs(key).uniform(size=tuple(shape)) * (maxval - minval) + minval).astype(_np.float32))
    DRAWS.append(("uniform", out))
    return out


def normal(key, shape, dtype=None):
    out = jnp._down(_rs(key).normal(size=tuple(shape)).astype(_np.float32))
    DRAWS.append(("normal", out))
    return out
