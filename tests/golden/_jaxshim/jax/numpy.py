"""numpy-backed stand-in for the subset of jax.numpy the reference's hot path uses (x64 disabled: every
64-bit result is truncated to 32 bits, like JAX's default).  Test infrastructure only."""
import numpy as _np

pi = _np.pi
inf = _np.inf
float32 = _np.float32
int32 = _np.int32
newaxis = None


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


_LOOP_OWNED = []      # stack of id-sets of arrays owned by the running lax.fori_loop (see lax.fori_loop)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        if _LOOP_OWNED and id(self.arr) in _LOOP_OWNED[-1]:      # carry of a lax.fori_loop: exclusively owned, update in place
            self.arr[self.idx] = v
            return self.arr
        out = _np.array(self.arr, copy=True).view(ndarray)
        out[self.idx] = v
        return out

    def add(self, v):
        out = _np.array(self.arr, copy=True).view(ndarray)
        _np.add.at(out, self.idx, v)
        return out


# jax.config.update("jax_enable_x64", True) stand-in: with X64 set, 64-bit types are kept (used only by the finite-difference
# goldens of make_reference_goldens.py, where the reference's own scan is evaluated in float64)
X64 = False


class ndarray(_np.ndarray):
    @property
    def at(self):
        return _At(self)

    # JAX type semantics (x64 disabled): no 64-bit types, and float32 (op) int32 -> float32 (numpy would give float64)
    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        arrs = [i.view(_np.ndarray) if isinstance(i, _np.ndarray) else i for i in inputs]
        any_float = any((isinstance(a, _np.ndarray) and a.dtype.kind == "f") or isinstance(a, (float, _np.floating))
                        for a in arrs)
        conv = []
        for a in arrs:
            if X64:
                pass
            elif isinstance(a, _np.ndarray):
                if a.dtype == _np.float64:
                    a = a.astype(_np.float32)
                elif a.dtype == _np.int64:
                    a = a.astype(_np.int32)
                if any_float and a.dtype.kind in "iu":
                    a = a.astype(_np.float32)
            elif isinstance(a, _np.float64):
                a = float(a)            # weak python scalar
            conv.append(a)
        return _down(getattr(ufunc, method)(*conv, **kwargs))

    def astype(self, dtype, *a, **k):
        dt = _np.dtype(dtype)
        if X64:
            pass
        elif dt == _np.int64:
            dt = _np.dtype(_np.int32)
        elif dt == _np.float64:
            dt = _np.dtype(_np.float32)
        return _np.ndarray.astype(self.view(_np.ndarray), dt, *a, **k).view(ndarray)

    # JAX arrays are immutable: `x += y` rebinds x to a new (broadcast) array and never mutates the operand
    def __iadd__(self, o):
        return _down(_np.add(self, o))

    def __isub__(self, o):
        return _down(_np.subtract(self, o))

    def __imul__(self, o):
        return _down(_np.multiply(self, o))

    def __itruediv__(self, o):
        return _down(_np.true_divide(self, o))


def _down(x):
    if isinstance(x, tuple):
        return tuple(_down(v) for v in x)
    if isinstance(x, list):
        return [_down(v) for v in x]
    if isinstance(x, _np.ndarray):
        if X64:
            return x.view(ndarray)
        if x.dtype == _np.float64:
            x = x.astype(_np.float32)
        elif x.dtype == _np.int64:
            x = x.astype(_np.int32)
        return x.view(ndarray)
    if X64:
        return x
    if isinstance(x, _np.float64):
        return _np.float32(x)
    if isinstance(x, _np.int64):
        return _np.int32(x)
    return x


def _wrap(fn):
    def inner(*a, **k):
        return _down(fn(*a, **k))
    inner.__name__ = getattr(fn, "__name__", "fn")
    return inner


def array(x, dtype=None):
    return _down(_np.array(x, dtype=dtype))


asarray = array


def nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
    # same positional signature as jnp.nan_to_num: the reference passes jnp.inf as `copy`
    return _down(_np.nan_to_num(x, copy=bool(copy), nan=nan, posinf=posinf, neginf=neginf))


def finfo(dt):
    return _np.finfo(_np.dtype(dt))


class _Linalg:
    @staticmethod
    def norm(x, ord=None, axis=None, keepdims=False):
        x = _np.asarray(x)
        return _down(_np.sqrt(_np.sum(x * x, axis=axis, keepdims=keepdims)))


linalg = _Linalg()


def __getattr__(name):
    fn = getattr(_np, name)
    if callable(fn) and not isinstance(fn, type):
        return _wrap(fn)
    return fn
