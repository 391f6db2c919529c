"""A few hundred lines of flax.linen semantics: dataclass modules, setup(), @compact, param scopes with Flax's
auto-naming (Dense_0, Dense_1, ...), nn.Dense and nn.scan -- enough to execute rnerf/models.py unchanged."""
import dataclasses
import functools
from typing import Any, Optional

import numpy as _np

import jax
from jax import numpy as jnp
from jax.nn import relu, sigmoid, softplus, tanh, softmax  # noqa: F401

_STACK = []          # modules whose method is currently executing
_CTX = {"init": False, "key": 0}


def _next_key():
    _CTX["key"] += 1
    return _np.array([17, _CTX["key"] * 2654435761 & 0x7FFFFFFF], dtype=_np.uint32)


def compact(fn):
    @functools.wraps(fn)
    def inner(self, *a, **k):
        return self._run(fn, a, k, compact=True)
    inner._is_compact = True
    return inner


class Module:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        ann = dict(cls.__dict__.get("__annotations__", {}))
        ann["name"] = Optional[str]
        ann["parent"] = Any
        cls.__annotations__ = ann
        cls.name = dataclasses.field(default=None, kw_only=True)
        cls.parent = dataclasses.field(default=None, kw_only=True, repr=False)
        dataclasses.dataclass(cls, eq=False, repr=False)
        call = cls.__dict__.get("__call__")
        if call is not None and not getattr(call, "_is_compact", False):
            @functools.wraps(call)
            def wrapped(self, *a, __call=call, **k):
                return self._run(__call, a, k, compact=False)
            cls.__call__ = wrapped

    def __post_init__(self):
        object.__setattr__(self, "_scope", None)
        object.__setattr__(self, "_setup_done", False)
        object.__setattr__(self, "_auto", {})
        if self.parent is None and _STACK:
            object.__setattr__(self, "parent", _STACK[-1])
            p = _STACK[-1]
            if p._in_compact and self.name is None:
                base = type(self).__name__
                i = p._auto.get(base, 0)
                p._auto[base] = i + 1
                object.__setattr__(self, "name", f"{base}_{i}")
        object.__setattr__(self, "_in_compact", False)
        object.__setattr__(self, "_in_setup", False)

    def __setattr__(self, key, value):
        if isinstance(value, Module) and getattr(self, "_in_setup", False) and value.name is None:
            object.__setattr__(value, "name", key)
            object.__setattr__(value, "parent", self)
        object.__setattr__(self, key, value)

    # ---- scope handling
    def _bind(self):
        if self._scope is not None:
            return
        if self.parent is None:
            raise RuntimeError("unbound top-level module: use .init / .apply")
        self.parent._bind()
        ps = self.parent._scope
        scope = ps if getattr(self, "_share_parent_scope", False) else ps.setdefault(self.name, {}) if _CTX["init"] else ps.get(self.name, {})
        object.__setattr__(self, "_scope", scope)

    def _ensure_setup(self):
        self._bind()
        if not self._setup_done:
            object.__setattr__(self, "_setup_done", True)
            if hasattr(self, "setup"):
                object.__setattr__(self, "_in_setup", True)
                _STACK.append(self)
                try:
                    self.setup()
                finally:
                    _STACK.pop()
                    object.__setattr__(self, "_in_setup", False)

    def _run(self, fn, a, k, compact):
        self._ensure_setup()
        _STACK.append(self)
        prev = self._in_compact
        object.__setattr__(self, "_in_compact", compact)
        if compact:
            object.__setattr__(self, "_auto", {})
        try:
            return fn(self, *a, **k)
        finally:
            object.__setattr__(self, "_in_compact", prev)
            _STACK.pop()

    def param(self, name, init_fn, *shape_args):
        self._bind()
        if name not in self._scope:
            if not _CTX["init"]:
                raise KeyError(f"missing parameter {name} in {self.name}")
            self._scope[name] = init_fn(_next_key(), *shape_args)
        return self._scope[name]

    # ---- public API
    def _fresh(self):
        kw = {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.name not in ("parent",)}
        return type(self)(**kw)

    def init(self, key, *a, method=None, **k):
        m = self._fresh()
        tree = {}
        object.__setattr__(m, "_scope", tree)
        _CTX["init"] = True
        try:
            fn = getattr(m, method.__name__) if method is not None else m
            m._ensure_setup()
            fn(*a, **k)
        finally:
            _CTX["init"] = False
        return {"params": tree}

    def apply(self, variables, *a, method=None, **k):
        m = self._fresh()
        object.__setattr__(m, "_scope", variables["params"])
        fn = getattr(m, method.__name__) if method is not None else m
        m._ensure_setup()
        _STACK.append(m)
        try:
            return fn(*a, **k)
        finally:
            _STACK.pop()

    @property
    def variables(self):
        return {"params": self._scope}


class Dense(Module):
    features: int
    kernel_init: Any = None
    use_bias: bool = True

    @compact
    def __call__(self, x):
        init = self.kernel_init or jax.nn.initializers.glorot_uniform()
        kernel = self.param("kernel", init, (x.shape[-1], self.features))
        y = jnp._down(_np.dot(_np.asarray(x), _np.asarray(kernel)))
        if self.use_bias:
            y = y + self.param("bias", jax.nn.initializers.zeros, (self.features,))
        return y


def scan(target, variable_broadcast=None, split_rngs=None, in_axes=0, out_axes=0):
    """nn.scan(ModuleClass, ...) -> class whose instances run the inner module over the leading axis of xs.
    The inner module shares the scanned module's parameter scope (Flax lifts params/<name>/... unchanged)."""
    class Scanned(Module):
        def setup(self):
            inner = target()
            object.__setattr__(inner, "name", "__scan_body__")
            object.__setattr__(inner, "parent", self)
            object.__setattr__(inner, "_share_parent_scope", True)
            object.__setattr__(self, "inner", inner)

        def __call__(self, carry, xs):
            outs = []
            for x in xs:
                carry, o = self.inner(carry, x)
                outs.append(_np.asarray(o))
            return carry, jnp._down(_np.stack(outs, axis=0))

    Scanned.__name__ = "Scan" + getattr(getattr(target, "func", target), "__name__", "Body")
    return Scanned
