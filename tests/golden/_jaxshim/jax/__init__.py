"""Minimal numpy-backed `jax` so that the reference's hot-path source files import and run unchanged.
TEST INFRASTRUCTURE ONLY (tests/golden/make_reference_goldens.py)."""
from . import numpy, lax, random, nn  # noqa: F401
from . import numpy as _jnp


def tree_map(fn, tree, *rest):
    if isinstance(tree, dict):
        return {k: tree_map(fn, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)) and not hasattr(tree, "_fields"):
        return type(tree)(tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    if hasattr(tree, "_fields"):
        return type(tree)(*[tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree)])
    return fn(tree, *rest)


class tree_util:
    tree_map = staticmethod(tree_map)


def jit(fn, *a, **k):
    return fn


def process_index():
    return 0


def process_count():
    return 1


def device_count():
    return 1


def local_device_count():
    return 1


def host_id():
    return 0


class _Scipy:
    class signal:
        pass

    class special:
        pass


scipy = _Scipy()
