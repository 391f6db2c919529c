import numpy as _np
from . import numpy as jnp


def stop_gradient(x):
    return x


def fori_loop(lo, hi, body, val):
    """The carry is copied once on entry and is then exclusively owned by the loop, so `.at[].set` inside the body may
    update it in place (same values as JAX's functional update; without this a 10 000-ray sample_pdf loop copies its
    [B,192,9] carry 30 000 times)."""
    import jax as _jax
    val = _jax.tree_map(lambda a: _np.array(a, copy=True).view(jnp.ndarray) if isinstance(a, _np.ndarray) else a, val)
    jnp._LOOP_OWNED.append({id(a) for a in ([val] if isinstance(val, _np.ndarray) else list(_leaves(val))) if isinstance(a, _np.ndarray)})
    try:
        for i in range(int(lo), int(hi)):
            val = body(i, val)
    finally:
        jnp._LOOP_OWNED.pop()
    return val


def _leaves(tree):
    if isinstance(tree, dict):
        for v in tree.values():
            yield from _leaves(v)
    elif isinstance(tree, (list, tuple)):
        for v in tree:
            yield from _leaves(v)
    else:
        yield tree


class Precision:
    HIGHEST = "highest"


def conv_dimension_numbers(lhs_shape, rhs_shape, spec):
    assert tuple(spec) == ("NHWDC", "HWDIO", "NHWDC")
    return spec


def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation, rhs_dilation, dn):
    """Only the call made by rnerf/ior_utils.py:354: 1 batch, 1 channel, stride 1, 'VALID' (a correlation)."""
    assert padding == "VALID" and tuple(window_strides) == (1, 1, 1)
    x = _np.asarray(lhs)[0, ..., 0]
    k = _np.asarray(rhs)[..., 0, 0]
    out_shape = tuple(x.shape[i] - k.shape[i] + 1 for i in range(3))
    out = _np.zeros(out_shape, dtype=_np.float32)
    for a in range(k.shape[0]):
        for b in range(k.shape[1]):
            for c in range(k.shape[2]):
                out += k[a, b, c] * x[a:a + out_shape[0], b:b + out_shape[1], c:c + out_shape[2]]
    return jnp._down(out[None, ..., None])
