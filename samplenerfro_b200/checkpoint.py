"""Flax-compatible checkpoints: `checkpoint_<step>` files holding the msgpack
serialisation `flax.serialization.to_bytes` produces, with the reference's tree names, so that weights trained by the
reference render through this path (train.py:322,424-427; eval.py:124-152).

Interoperability is in the PARAMETER direction: files written by the reference are read here (parameters + step), and
files written here can be read by the reference wherever it restores with target=None and picks sub-trees
(`checkpoints.restore_checkpoint(dir, None)["params"]["params"][...]`, eval.py:124-152, train.py:286-310).  The
reference's `restore_checkpoint(stage_dir, state)` (train.py:322) runs `from_state_dict` against its
optax.multi_transform state and would reject the "arena_adam" optimiser state stored here: resuming THIS library's
training run inside the reference's train.py is not supported (and vice versa the reference's optimiser state is skipped).

Wire format (flax/serialization.py): a msgpack map; every ndarray is ExtType(1, packb((shape, dtype.name, bytes)));
numpy scalars are ExtType(3, same tuple).  The reference's state dict is
    {"step": int, "params": {"params": {coarse_mlp, fine_mlp, bkgd_mlp, path_sampler}}, "opt_state": ...}
(`pretrain["params"]["params"][...]`, eval.py:128-131).  The optimiser state of the reference (optax.multi_transform)
is not interpreted; this module stores its own Adam moments under "opt_state"/"arena_adam" and restores them when
present.  Needs only `msgpack` + numpy: no jax/flax.
"""
from __future__ import annotations

import os
import re
from typing import Any, Dict, Optional

import msgpack
import numpy as np
import torch

_EXT_NDARRAY, _EXT_NPSCALAR = 1, 3
PREFIX = "checkpoint_"


def _pack_array(a: np.ndarray) -> bytes:
    a = np.asarray(a, order="C")           # (ascontiguousarray would promote 0-d arrays to 1-d)
    return msgpack.packb((list(a.shape), a.dtype.name, a.tobytes("C")), use_bin_type=True)


def _default(o):
    if isinstance(o, torch.Tensor):
        o = o.detach().cpu().numpy()
    if isinstance(o, np.ndarray):
        return msgpack.ExtType(_EXT_NDARRAY, _pack_array(o))
    if isinstance(o, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, _pack_array(np.asarray(o)))
    raise TypeError(f"cannot serialise {type(o)}")


def _ext_hook(code, data):
    if code in (_EXT_NDARRAY, _EXT_NPSCALAR):
        shape, dtype, buf = msgpack.unpackb(data, raw=False)
        arr = np.frombuffer(buf, dtype=np.dtype(dtype)).reshape(shape).copy()
        return arr if code == _EXT_NDARRAY else arr[()]
    return msgpack.ExtType(code, data)


def to_bytes(tree: Dict) -> bytes:
    """flax.serialization.to_bytes for nested dicts of arrays / scalars."""
    return msgpack.packb(tree, default=_default, strict_types=True, use_bin_type=True)


def from_bytes(data: bytes) -> Dict:
    """flax.serialization.msgpack_restore: nested dicts with numpy leaves."""
    return msgpack.unpackb(data, ext_hook=_ext_hook, raw=False, strict_map_key=False)


def _to_numpy_tree(t):
    if isinstance(t, dict):
        return {k: _to_numpy_tree(v) for k, v in t.items()}
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t


def state_dict(state) -> Dict:
    """The dict `checkpoints.save_checkpoint` would write for a train.TrainState."""
    d = {"step": np.asarray(int(state.step), dtype=np.int64), "params": _to_numpy_tree(state.params)}
    if getattr(state, "arena", None) is not None and getattr(state, "opt", None) is not None:
        d["opt_state"] = {"arena_adam": {"count": np.asarray(int(state.opt.count), dtype=np.int64),
                                         "mu": state.opt.mu.detach().cpu().numpy(), "nu": state.opt.nu.detach().cpu().numpy()}}
    return d


def _natural_key(name: str):
    return [int(s) if s.isdigit() else s for s in re.split(r"(\d+)", name)]


def latest_checkpoint(ckpt_dir: str, prefix: str = PREFIX) -> Optional[str]:
    if not os.path.isdir(ckpt_dir):
        return None
    files = [f for f in os.listdir(ckpt_dir) if f.startswith(prefix) and not f.endswith("tmp")]
    return os.path.join(ckpt_dir, sorted(files, key=_natural_key)[-1]) if files else None


def save_checkpoint(ckpt_dir: str, state, step: int, prefix: str = PREFIX, keep: int = 100) -> str:
    """flax.training.checkpoints.save_checkpoint (train.py:424-427): atomic write of `<prefix><step>`, keep the newest
    `keep` files."""
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, f"{prefix}{int(step)}")
    tmp = path + "tmp"
    payload = to_bytes(state if isinstance(state, dict) else state_dict(state))
    with open(tmp, "wb") as f:
        f.write(payload)
    os.replace(tmp, path)
    files = sorted([f for f in os.listdir(ckpt_dir) if f.startswith(prefix) and not f.endswith("tmp")], key=_natural_key)
    for old in files[:-keep] if keep > 0 else []:
        os.remove(os.path.join(ckpt_dir, old))
    return path


def load_params_into(variables: Dict, loaded: Dict, names=None) -> None:
    """Copy loaded["params"]["params"][name] into the live variables tree in place (eval.py:128-131, 146-151): the
    leaves keep their storage (e.g. ParamArena views), shapes must match."""
    src = loaded["params"]["params"] if "params" in loaded.get("params", {}) else loaded["params"]
    names = list(src.keys()) if names is None else names

    def copy(dst, s, where):
        for k, v in s.items():
            if k not in dst:
                raise KeyError(f"checkpoint has {where}/{k}, the model does not")
            if isinstance(v, dict):
                copy(dst[k], v, f"{where}/{k}")
            else:
                a = torch.from_numpy(np.asarray(v))
                if tuple(a.shape) != tuple(dst[k].shape):
                    raise ValueError(f"{where}/{k}: checkpoint shape {tuple(a.shape)} != model shape {tuple(dst[k].shape)}")
                with torch.no_grad():
                    dst[k].copy_(a.to(dst[k].device, dst[k].dtype))

    for n in names:
        copy(variables["params"][n], src[n], n)


def restore_checkpoint(ckpt_dir_or_file: str, target=None, prefix: str = PREFIX):
    """flax.training.checkpoints.restore_checkpoint.  target=None returns the raw state dict (eval.py:126); a
    train.TrainState target is updated in place (parameters, step, and this library's Adam moments when the file
    has them) and returned; if nothing is found the target is returned unchanged (train.py:322)."""
    path = ckpt_dir_or_file if os.path.isfile(ckpt_dir_or_file) else latest_checkpoint(ckpt_dir_or_file, prefix)
    if path is None:
        return target
    with open(path, "rb") as f:
        loaded = from_bytes(f.read())
    if target is None:
        return loaded
    load_params_into(target.params, loaded)
    target.step = int(loaded["step"])
    adam = loaded.get("opt_state", {}).get("arena_adam") if isinstance(loaded.get("opt_state"), dict) else None
    if adam is not None and getattr(target, "opt", None) is not None:
        n_file, n_here = int(np.asarray(adam["mu"]).size), int(target.opt.mu.numel())
        if n_file != n_here:
            # the moments cover the TRAINABLE buckets of the stage that wrote the file (radiance: 3 MLPs; "all": + so3_mlp);
            # restoring into another stage keeps the parameters and starts the optimiser fresh, like the reference does when
            # it loads radiance weights into the "all" stage by sub-tree (train.py:286-330 / eval.py:124-152)
            import warnings
            warnings.warn(f"checkpoint Adam moments cover {n_file} values, this stage trains {n_here}: optimiser state not restored")
            return target
        with torch.no_grad():
            target.opt.mu.copy_(torch.from_numpy(adam["mu"]).to(target.opt.mu.device))
            target.opt.nu.copy_(torch.from_numpy(adam["nu"]).to(target.opt.nu.device))
        target.opt.count = int(adam["count"])
    return target
