"""Thin torch-tensor wrappers over the C ABI (include/rnerf_b200.h).

PyTorch owns the device buffers and the stream; every op below is one call into librnerf_b200.so.
There is no alternative implementation: without the library (or a CUDA tensor) these raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, NamedTuple, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Dbl3, Int3, check

PATH_STRIDE = 12
PATH_STRIDE_COMPACT = 8


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.RnerfError(f"{name}: expected a CUDA tensor (the rendering path has no CPU implementation)")
    if t.dtype != dtype:
        raise _lib.RnerfError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.RnerfError(f"{name}: expected a contiguous tensor")
    return t


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _geom(ndim, nmin, nmax):
    return Int3(*[int(v) for v in ndim]), Dbl3(*[float(v) for v in nmin]), Dbl3(*[float(v) for v in nmax])


# ---------------------------------------------------------------- grid (a1, a2, a3)
def grid_blur(n: torch.Tensor, ndim: Sequence[int], ws: int, sigma: float) -> torch.Tensor:
    """conv3d_normal (rnerf/ior_utils.py:327-363).  n: [G^3] or [G^3,1] fp32 -> [G^3,1]."""
    n = _chk(n.reshape(-1), "n")
    out = torch.empty_like(n)
    nd = Int3(*[int(v) for v in ndim])
    check(_lib.load().rnerf_grid_blur(_p(n), _p(out), nd, int(ws), float(sigma), _stream()), "rnerf_grid_blur")
    return out.reshape(-1, 1)


def grid_table(n: torch.Tensor, ndim, nmin, nmax, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """VoxMLP.setup (rnerf/ior_utils.py:161): [G^3,4] = (n, grad n)."""
    n = _chk(n.reshape(-1), "n")
    table = torch.empty(n.numel(), 4, device=n.device, dtype=torch.float32) if out is None else _chk(out, "out")
    assert table.numel() == 4 * n.numel()
    nd, lo, hi = _geom(ndim, nmin, nmax)
    check(_lib.load().rnerf_grid_table(_p(n), nd, lo, hi, _p(table), _stream()), "rnerf_grid_table")
    return table


def grid_lookup(table: torch.Tensor, ndim, nmin, nmax, pts: torch.Tensor) -> torch.Tensor:
    """VoxMLP._linear3 (rnerf/ior_utils.py:188-223).  pts [N,3] -> [N,4]."""
    _chk(table, "table"); pts = _chk(pts, "pts")
    out = torch.empty(pts.shape[0], 4, device=pts.device, dtype=torch.float32)
    nd, lo, hi = _geom(ndim, nmin, nmax)
    check(_lib.load().rnerf_grid_lookup(_p(table), nd, lo, hi, _p(pts), pts.shape[0], _p(out), _stream()),
          "rnerf_grid_lookup")
    return out


# ---------------------------------------------------------------- march (a5, a6) / select (a7)
def grid_bricks(table: torch.Tensor, ndim, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Brick map of a (n, grad n) table: per 8^3-voxel brick the common n, or NaN if the brick is not homogeneous."""
    _chk(table, "table")
    lib = _lib.load()
    nd = Int3(*[int(v) for v in ndim])
    n_bricks = int(lib.rnerf_grid_brick_count(nd))
    bricks = torch.empty(n_bricks, device=table.device, dtype=torch.float32) if out is None else _chk(out, "out")
    assert bricks.numel() == n_bricks
    check(lib.rnerf_grid_bricks(_p(table), nd, _p(bricks), _stream()), "rnerf_grid_bricks")
    return bricks


class BentPath(NamedTuple):
    """The marched path: interleaved records [B,S,12] (full) or [B,S,8] (compact, no idx_grad) and the dense
    ray_dist column [B,S] (or None).  Layout: include/rnerf_b200.h."""
    rec: torch.Tensor
    t: Optional[torch.Tensor]

    @property
    def compact(self) -> bool:
        return self.rec.shape[-1] == PATH_STRIDE_COMPACT


def so3_saved_buffer(n_rays: int, n_steps: int, device, max_bytes: int = 24 << 30) -> Optional[torch.Tensor]:
    """Uninitialised buffer for the hidden activations of the so3_mlp evaluations of a training forward (2 KB per (ray, step)
    slot; only evaluated slots are ever written or read), or None when the launch would not run the kernel that writes it
    (a full-frame launch, RNERF_SO3_TC=0) or the buffer would exceed `max_bytes`."""
    n = int(_lib.load().rnerf_so3_saved_floats(int(n_rays), int(n_steps)))
    if n == 0 or 4 * n > max_bytes:
        return None
    return torch.empty(n, device=device, dtype=torch.float32)


def so3_pack(p: Dict) -> torch.Tensor:
    """so3_mlp parameters (model_utils.MLP 60 -> 128 x4 (+60 after layer 2) -> 3, rnerf/ior_utils.py:147-152) as the flat
    fp32 image rnerf_march_all_fwd reads: the 5 kernels, then the 5 biases."""
    ks = [_chk(p[f"Dense_{i}"]["kernel"], f"so3 Dense_{i}.kernel") for i in range(5)]
    bs = [_chk(p[f"Dense_{i}"]["bias"], f"so3 Dense_{i}.bias") for i in range(5)]
    expect = [(60, 128), (128, 128), (128, 128), (188, 128), (128, 3)]
    for i, (k, e) in enumerate(zip(ks, expect)):
        if tuple(k.shape) != e:
            raise _lib.RnerfError(f"so3 Dense_{i}.kernel has shape {tuple(k.shape)}, expected {e}")
    w = torch.cat([k.reshape(-1) for k in ks] + [b.reshape(-1) for b in bs]).contiguous()
    assert w.numel() == _lib.load().rnerf_so3_weight_floats()
    return w


def _window_args(window):
    """so3 window as the (host double[10] or None, device fp32[10] or None) pair of the C ABI: a CUDA tensor is passed by
    pointer (read at run time -- what a captured graph needs), anything else by value."""
    if isinstance(window, torch.Tensor) and window.is_cuda:
        _chk(window, "so3 window")
        assert window.numel() == 10
        return None, _p(window)
    vals = [float(v) for v in window]
    assert len(vals) == 10
    return (C.c_double * 10)(*vals), None


def march(table, ndim, nmin, nmax, origins, viewdirs, near: float, far: float, n_steps: int,
          out: Optional[BentPath] = None, bricks: Optional[torch.Tensor] = None, compact: bool = False,
          t_col: bool = True, so3: Optional[Tuple[torch.Tensor, Sequence[float]]] = None,
          so3_tc: Optional[torch.Tensor] = None, so3_saved: Optional[torch.Tensor] = None) -> BentPath:
    """PathSampler.__call__ (rnerf/eikonal_utils.py:101-124).  Returns the BentPath.
    `so3_saved` = so3_saved_buffer(B, n_steps) (training): the forward leaves the hidden activations of every so3_mlp
    evaluation there and march_all_bwd reads them back instead of recomputing them.
    `bricks` (from grid_bricks) lets the kernel skip the gathers in homogeneous space; results are bit-identical.
    `compact` drops idx_grad from the records (8 instead of 12 floats per step).
    `so3` = (so3_pack(...) weights, 10 window values): the "all" stage, where every step rotates grad n by the so3_mlp
    prediction wherever |grad n| > 1e-3 (rnerf/eikonal_utils.py:34-35); None = radiance stage.
    `so3_tc` = so3_tc_pack(weights): lets full-frame launches with compact records run so3_mlp on the tensor pipe."""
    _chk(table, "table"); origins = _chk(origins, "origins"); viewdirs = _chk(viewdirs, "viewdirs")
    if bricks is not None:
        _chk(bricks, "bricks")
    B = origins.shape[0]
    W = PATH_STRIDE_COMPACT if compact else PATH_STRIDE
    if out is None:
        rec = torch.empty(B, n_steps, W, device=origins.device, dtype=torch.float32)
        out = BentPath(rec, torch.empty(B, n_steps, device=origins.device, dtype=torch.float32) if t_col else None)
    else:
        _chk(out.rec, "out.rec")
        assert out.rec.shape == (B, n_steps, W)
        if out.t is not None:
            _chk(out.t, "out.t")
            assert out.t.shape == (B, n_steps)
    nd, lo, hi = _geom(ndim, nmin, nmax)
    if so3 is not None:
        w, window = so3
        _chk(w, "so3 weights")
        win, win_dev = _window_args(window)
        check(_lib.load().rnerf_march_all_fwd(_p(table), _p(bricks), nd, lo, hi, _p(origins), _p(viewdirs), B, float(near),
                                              float(far), int(n_steps), W, _p(w), win, win_dev, _p(so3_tc),
                                              _p(None if so3_saved is None else _chk(so3_saved, "so3_saved")), _p(out.rec), _p(out.t),
                                              _stream()),
              "rnerf_march_all_fwd")
        return out
    check(_lib.load().rnerf_march_fwd(_p(table), _p(bricks), nd, lo, hi, _p(origins), _p(viewdirs), B, float(near),
                                      float(far), int(n_steps), W, _p(out.rec), _p(out.t), _stream()), "rnerf_march_fwd")
    return out


def _rec(path) -> torch.Tensor:
    rec = path.rec if isinstance(path, BentPath) else path
    _chk(rec, "path")
    if rec.dim() != 3 or rec.shape[-1] not in (PATH_STRIDE, PATH_STRIDE_COMPACT):
        raise _lib.RnerfError(f"path records must be [B,S,12] or [B,S,8], got {tuple(rec.shape)}")
    return rec


def path_dirs(path) -> torch.Tensor:
    """ray_dir [B,S,3]: safe_l2_normalize of the direction state stored in every record."""
    rec = _rec(path)
    B, S, W = rec.shape
    out = torch.empty(B, S, 3, device=rec.device, dtype=torch.float32)
    check(_lib.load().rnerf_path_dirs(_p(rec), W, B, S, _p(out), _stream()), "rnerf_path_dirs")
    return out


def path_views(path):
    """(ray_pos, ray_dir, ray_dist, idx_data, idx_grad) as PathSampler returns them: strided views of the records,
    except ray_dir which is normalised on demand (the path stores the un-normalised direction state).  idx_grad is
    None for compact records."""
    rec = _rec(path)
    return rec[..., 0:3], path_dirs(rec), rec[..., 3], rec[..., 7:8], (rec[..., 8:11] if rec.shape[-1] == PATH_STRIDE else None)


def select(path, jitter: torch.Tensor, want_grad: bool = False):
    """rnerf/models.py:243-247: gather the coarse samples at march-step indices `jitter` (int32 [Nc])."""
    rec = _rec(path); jitter = _chk(jitter, "jitter", torch.int32)
    B, S, W = rec.shape
    Nc = jitter.numel()
    dev = rec.device
    pos = torch.empty(B, Nc, 3, device=dev); dirs = torch.empty(B, Nc, 3, device=dev); t = torch.empty(B, Nc, device=dev)
    grad = torch.empty(B, Nc, 3, device=dev) if want_grad else None
    check(_lib.load().rnerf_select(_p(rec), W, B, S, _p(jitter), Nc, _p(pos), _p(dirs), _p(t), _p(grad), _stream()),
          "rnerf_select")
    return pos, dirs, t, grad


def so3_unpack_views(g: torch.Tensor):
    """Views of a flat so3 image (weights or gradients) in Flax order: [K0, b0, ..., K4, b4]."""
    shapes = [(60, 128), (128, 128), (128, 128), (188, 128), (128, 3)]
    ks, off = [], 0
    for sh in shapes:
        n = sh[0] * sh[1]
        ks.append(g[off:off + n].view(sh)); off += n
    out = []
    for k, n in zip(ks, (128, 128, 128, 128, 3)):
        out += [k, g[off:off + n]]; off += n
    return out


def so3_predict(w: torch.Tensor, window: Sequence[float], pts: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
    """VoxMLP.wrapper_grad_mlp (rnerf/ior_utils.py:225-267): rodrigues(so3_mlp(annealed_pos_enc(pts)), cond) -> [N,3]."""
    pts = _chk(pts.reshape(-1, 3).contiguous(), "pts"); cond = _chk(cond.reshape(-1, 3).contiguous(), "cond")
    assert pts.shape == cond.shape
    pred = torch.empty_like(pts)
    win, win_dev = _window_args(window)
    check(_lib.load().rnerf_so3_predict(_p(_chk(w, "so3 weights")), win, win_dev, _p(pts), _p(cond), pts.shape[0], _p(pred), _stream()),
          "rnerf_so3_predict")
    return pred


def so3_tc_pack(w: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """hi / lo TF32 halves of the so3 kernels as the pre-swizzled chunks the tensor-pipe evaluator streams (so3_tc.cuh)."""
    lib = _lib.load()
    if out is None:
        out = torch.empty(lib.rnerf_so3_tc_packed_bytes(), device=w.device, dtype=torch.uint8)
    check(lib.rnerf_so3_tc_pack(_p(_chk(w, "so3 weights")), _p(out), _stream()), "rnerf_so3_tc_pack")
    return out


def so3_predict_tc(packed: torch.Tensor, w: torch.Tensor, window, pts: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
    """so3_predict through the tensor-pipe evaluator (tcgen05 kind::tf32, 3xTF32 split)."""
    pts = _chk(pts.reshape(-1, 3).contiguous(), "pts"); cond = _chk(cond.reshape(-1, 3).contiguous(), "cond")
    assert pts.shape == cond.shape
    pred = torch.empty_like(pts)
    win, win_dev = _window_args(window)
    check(_lib.load().rnerf_so3_predict_tc(_p(_chk(packed, "so3 tc image", torch.uint8)), _p(_chk(w, "so3 weights")), win, win_dev,
                                           _p(pts), _p(cond), pts.shape[0], _p(pred), _stream()), "rnerf_so3_predict_tc")
    return pred


def so3_transpose(w: torch.Tensor) -> torch.Tensor:
    """T_l[out][in] images of the four hidden so3 kernels (the adjoint's input-gradient GEMMs stream them row by row)."""
    lib = _lib.load()
    wt = torch.empty(lib.rnerf_so3_transposed_floats(), device=w.device, dtype=torch.float32)
    check(lib.rnerf_so3_transpose(_p(_chk(w, "so3 weights")), _p(wt), _stream()), "rnerf_so3_transpose")
    return wt


def march_all_bwd(table, ndim, nmin, nmax, path, near: float, far: float, jitter: torch.Tensor, d_pos_c: torch.Tensor,
                  d_dir_c: torch.Tensor, so3: Optional[Tuple[torch.Tensor, Sequence[float]]], bricks: Optional[torch.Tensor] = None,
                  g_so3: Optional[torch.Tensor] = None, want_ray_grads: bool = False, d_table: Optional[torch.Tensor] = None,
                  so3_saved: Optional[torch.Tensor] = None):
    """Reverse sweep of the "all"-stage scan (rnerf/eikonal_utils.py:30-49,75-82 under jax.value_and_grad, train.py:164):
    from the loss gradients of the coarse samples, d_pos_c / d_dir_c [B,Nc,3] at march steps `jitter` (strictly
    increasing), to the gradient of so3_mlp.  Returns (g_so3 [so3 image layout, accumulated into when given],
    d_origins, d_viewdirs [B,3] or None).
    Extension (no reference counterpart): `d_table` [G^3,4], when given, is ACCUMULATED with the gradient wrt the (n, grad n)
    table (take it on to the n-grid with grid_table_bwd); `so3=None` sweeps a radiance-stage path (g_so3 is then None)."""
    rec = _rec(path); jitter = _chk(jitter, "jitter", torch.int32)
    B, S, W = rec.shape
    Nc = jitter.numel()
    w, window = so3 if so3 is not None else (None, None)
    _chk(table, "table")
    if d_table is not None:
        _chk(d_table, "d_table")
        assert d_table.numel() == table.numel()
    if w is None:
        d_pos_c = _chk(d_pos_c.contiguous(), "d_pos_c"); d_dir_c = _chk(d_dir_c.contiguous(), "d_dir_c")
        assert d_pos_c.shape == (B, Nc, 3) and d_dir_c.shape == (B, Nc, 3)
        d_o = torch.empty(B, 3, device=rec.device) if want_ray_grads else None
        d_d = torch.empty(B, 3, device=rec.device) if want_ray_grads else None
        nd, lo, hi = _geom(ndim, nmin, nmax)
        check(_lib.load().rnerf_march_all_bwd(_p(table), _p(bricks), nd, lo, hi, _p(rec), W, B, float(near), float(far), S,
                                              _p(jitter), Nc, _p(d_pos_c), _p(d_dir_c), None, None, None, None, None, None, _p(d_o),
                                              _p(d_d), _p(d_table), _stream()), "rnerf_march_all_bwd")
        return None, d_o, d_d
    _chk(w, "so3 weights")
    d_pos_c = _chk(d_pos_c.contiguous(), "d_pos_c"); d_dir_c = _chk(d_dir_c.contiguous(), "d_dir_c")
    assert d_pos_c.shape == (B, Nc, 3) and d_dir_c.shape == (B, Nc, 3)
    if bricks is not None:
        _chk(bricks, "bricks")
    g = torch.zeros_like(w) if g_so3 is None else _chk(g_so3, "g_so3")
    assert g.numel() == w.numel()
    wt = so3_transpose(w)
    d_o = torch.empty(B, 3, device=rec.device) if want_ray_grads else None
    d_d = torch.empty(B, 3, device=rec.device) if want_ray_grads else None
    nd, lo, hi = _geom(ndim, nmin, nmax)
    win, win_dev = _window_args(window)
    check(_lib.load().rnerf_march_all_bwd(_p(table), _p(bricks), nd, lo, hi, _p(rec), W, B, float(near), float(far), S,
                                          _p(jitter), Nc, _p(d_pos_c), _p(d_dir_c), _p(w), _p(wt), win, win_dev,
                                          _p(None if so3_saved is None else _chk(so3_saved, "so3_saved")), _p(g), _p(d_o),
                                          _p(d_d), _p(d_table), _stream()), "rnerf_march_all_bwd")
    return g, d_o, d_d


def grid_table_bwd(d_table: torch.Tensor, ndim, nmin, nmax) -> torch.Tensor:
    """Adjoint of grid_table (extension: gradients of a learned IoR grid): d_table [G^3,4] -> d_n [G^3]."""
    _chk(d_table, "d_table")
    d_n = torch.empty(d_table.numel() // 4, device=d_table.device, dtype=torch.float32)
    nd, lo, hi = _geom(ndim, nmin, nmax)
    check(_lib.load().rnerf_grid_table_bwd(_p(d_table), nd, lo, hi, _p(d_n), _stream()), "rnerf_grid_table_bwd")
    return d_n


# ---------------------------------------------------------------- encoding-fused radiance MLP (a8, a9)
def nerf_mlp_layer_list(p: Dict) -> Tuple[list, list]:
    ks, bs = [], []
    for i in range(12):
        d = p[f"Dense_{i}"]
        ks.append(_chk(d["kernel"], f"Dense_{i}.kernel")); bs.append(_chk(d["bias"], f"Dense_{i}.bias"))
    expect = [(63, 256)] + [(256, 256)] * 4 + [(319, 256)] + [(256, 256)] * 2 + [(256, 1), (256, 256), (283, 128), (128, 3)]
    for i, (k, e) in enumerate(zip(ks, expect)):
        if tuple(k.shape) != e:
            raise _lib.RnerfError(f"Dense_{i}.kernel has shape {tuple(k.shape)}, the sm_100a kernel is built for {e} "
                                  "(net_depth=8, net_width=256, skip_layer=4, condition 1x128, deg 10/4)")
    return ks, bs


def encmlp_pack(p: Dict, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pack the 12 Flax Dense layers of a NerfMLP into the pre-swizzled bf16 image the kernel streams."""
    ks, bs = nerf_mlp_layer_list(p)
    lib = _lib.load()
    nbytes = lib.rnerf_encmlp_packed_bytes()
    if out is None:
        out = torch.empty(nbytes, device=ks[0].device, dtype=torch.uint8)
    kp = (C.c_void_p * 12)(*[k.data_ptr() for k in ks])
    bp = (C.c_void_p * 12)(*[b.data_ptr() for b in bs])
    check(lib.rnerf_encmlp_pack(kp, bp, _p(out), _stream()), "rnerf_encmlp_pack")
    return out


def encmlp_fwd(packed: torch.Tensor, pos: torch.Tensor, dirs: torch.Tensor, debug_layers: bool = False):
    """pos_enc + NerfMLP.  pos/dirs [M,3] (or [B,Ns,3]) -> raw [M,4] = (raw_rgb, raw_sigma)."""
    _chk(packed, "packed", torch.uint8)
    pos = _chk(pos, "pos").reshape(-1, 3); dirs = _chk(dirs, "dirs").reshape(-1, 3)
    M = pos.shape[0]
    raw = torch.empty(M, 4, device=pos.device, dtype=torch.float32)
    lib = _lib.load()
    if debug_layers:
        layers = torch.zeros(10, M, 256, device=pos.device, dtype=torch.bfloat16)
        check(lib.rnerf_encmlp_fwd_debug(_p(packed), _p(pos), _p(dirs), M, _p(raw), _p(layers), _stream()),
              "rnerf_encmlp_fwd_debug")
        return raw, layers
    check(lib.rnerf_encmlp_fwd(_p(packed), _p(pos), _p(dirs), M, _p(raw), _stream()), "rnerf_encmlp_fwd")
    return raw


# ---------------------------------------------------------------- background MLP (a10)
def bkgd_pack(p: Dict) -> torch.Tensor:
    ks = [_chk(p[f"Dense_{i}"]["kernel"], f"bkgd Dense_{i}.kernel") for i in range(5)]
    bs = [_chk(p[f"Dense_{i}"]["bias"], f"bkgd Dense_{i}.bias") for i in range(5)]
    expect = [(27, 128), (128, 128), (128, 128), (155, 128), (128, 3)]
    for i, (k, e) in enumerate(zip(ks, expect)):
        if tuple(k.shape) != e:
            raise _lib.RnerfError(f"bkgd Dense_{i}.kernel has shape {tuple(k.shape)}, expected {e}")
    w = torch.cat([k.reshape(-1) for k in ks] + [b.reshape(-1) for b in bs]).contiguous()
    assert w.numel() == _lib.load().rnerf_bkgd_weight_floats()
    return w


def bkgd_mlp_fwd(w: torch.Tensor, dirs: torch.Tensor, n_rays: int, stride_floats: int = 3, offset_floats: int = 0):
    """bkgd_mlp(pos_enc(dir)) -> raw [B,3].  `dirs` may be a larger array read with a ray stride (e.g. the
    last coarse sample of a [B,Nc,3] array: offset (Nc-1)*3, stride Nc*3)."""
    _chk(w, "w"); _chk(dirs, "dirs")
    out = torch.empty(n_rays, 3, device=dirs.device, dtype=torch.float32)
    ptr = C.c_void_p(dirs.data_ptr() + 4 * offset_floats)
    check(_lib.load().rnerf_bkgd_mlp_fwd(_p(w), ptr, n_rays, stride_floats, _p(out), _stream()), "rnerf_bkgd_mlp_fwd")
    return out


def bkgd_tc_pack(w: torch.Tensor):
    """Background weights (bkgd_pack layout) -> (the same zero-padded into so3_mlp's layout, its tensor-pipe image): the two
    networks have the same shape, so the so3 evaluator runs the background MLP of a whole frame (bkgd_mlp_fwd_tc)."""
    _chk(w, "w")
    k0, k1, k2, k3, k4 = 27 * 128, 128 * 128, 128 * 128, 155 * 128, 128 * 3
    o = [0, k0, k0 + k1, k0 + k1 + k2, k0 + k1 + k2 + k3, k0 + k1 + k2 + k3 + k4]
    z = lambda rows: torch.zeros(rows * 128, device=w.device, dtype=torch.float32)
    img = torch.cat([w[o[0]:o[1]], z(60 - 27), w[o[1]:o[2]], w[o[2]:o[3]], w[o[3]:o[4]], z(188 - 155), w[o[4]:o[5]], w[o[5]:]]).contiguous()
    assert img.numel() == _lib.load().rnerf_so3_weight_floats()
    return img, so3_tc_pack(img)


def bkgd_mlp_fwd_tc(tc, dirs: torch.Tensor, n_rays: int, stride_floats: int = 3, offset_floats: int = 0):
    """bkgd_mlp_fwd on the tensor pipe; tc = bkgd_tc_pack(w)."""
    img, packed = tc
    _chk(dirs, "dirs")
    out = torch.empty(n_rays, 3, device=dirs.device, dtype=torch.float32)
    ptr = C.c_void_p(dirs.data_ptr() + 4 * offset_floats)
    check(_lib.load().rnerf_bkgd_mlp_fwd_tc(_p(packed), _p(img), ptr, n_rays, stride_floats, _p(out), _stream()),
          "rnerf_bkgd_mlp_fwd_tc")
    return out


# ---------------------------------------------------------------- compositing (a11, a12)
def composite_fwd(raw, t, dirs, bkgd_raw=None, mask=None, white_bkgd=False, rgb_padding=0.001, sigma_bias=-1.0,
                  want_weights=True, want_alpha=False):
    """activations + volumetric_rendering (rnerf/models.py:334-349, rnerf/model_utils.py:247-309).
    raw [B,Ns,4], t [B,Ns], dirs [B,Ns,3] -> dict(comp_rgb, distance, acc, weights, alpha, trans, trans_rgb_bkgd)."""
    raw = _chk(raw, "raw"); t = _chk(t, "t"); dirs = _chk(dirs, "dirs")
    B, Ns = t.shape
    dev = t.device
    o = {"comp_rgb": torch.empty(B, 3, device=dev), "distance": torch.empty(B, device=dev),
         "acc": torch.empty(B, device=dev), "trans": torch.empty(B, 1, device=dev),
         "trans_rgb_bkgd": torch.empty(B, 3, device=dev),
         "weights": torch.empty(B, Ns, device=dev) if want_weights else None,
         "alpha": torch.empty(B, Ns, device=dev) if want_alpha else None}
    if bkgd_raw is not None:
        _chk(bkgd_raw, "bkgd_raw")
    if mask is not None:
        _chk(mask, "mask")
    check(_lib.load().rnerf_composite_fwd(_p(raw), _p(t), _p(dirs), _p(bkgd_raw), _p(mask), B, Ns, int(white_bkgd),
                                          float(rgb_padding), float(sigma_bias), _p(o["comp_rgb"]), _p(o["distance"]),
                                          _p(o["acc"]), _p(o["weights"]), _p(o["alpha"]), _p(o["trans"]),
                                          _p(o["trans_rgb_bkgd"]), _stream()), "rnerf_composite_fwd")
    return o


def composite_bwd(raw, t, dirs, bkgd_raw, mask, d_comp_rgb, d_trans, d_trb, white_bkgd=False, rgb_padding=0.001,
                  sigma_bias=-1.0):
    raw = _chk(raw, "raw"); t = _chk(t, "t"); dirs = _chk(dirs, "dirs")
    B, Ns = t.shape
    d_raw = torch.empty(B, Ns, 4, device=t.device)
    d_bk = torch.empty(B, 3, device=t.device) if bkgd_raw is not None else None
    for nm, x in (("d_comp_rgb", d_comp_rgb), ("d_trans", d_trans), ("d_trb", d_trb)):
        if x is not None:
            _chk(x, nm)
    check(_lib.load().rnerf_composite_bwd(_p(raw), _p(t), _p(dirs), _p(bkgd_raw), _p(mask), B, Ns, int(white_bkgd),
                                          float(rgb_padding), float(sigma_bias), _p(d_comp_rgb), _p(d_trans), _p(d_trb),
                                          _p(d_raw), _p(d_bk), _stream()), "rnerf_composite_bwd")
    return d_raw, d_bk


# ---------------------------------------------------------------- resampling (a13, a14)
def resample(path, t_c, weights_c, u, n_fine: int, want_grad: bool = False):
    """sorted_piecewise_constant_pdf + sample_pdf (rnerf/model_utils.py:312-435).
    u: [Nf] (shared) or [B,Nf] sorted CDF positions.  Returns t_f [B,Nc+Nf], pos_f, dir_f, grad_f."""
    rec = _rec(path); t_c = _chk(t_c, "t_c"); weights_c = _chk(weights_c, "weights_c"); u = _chk(u, "u")
    tcol = path.t if isinstance(path, BentPath) else None
    if tcol is not None:
        _chk(tcol, "path.t")
    B, S, W = rec.shape
    Nc = t_c.shape[1]
    per_ray = 1 if u.dim() == 2 else 0
    assert u.shape[-1] == n_fine and (not per_ray or u.shape[0] == B)
    Nt = Nc + n_fine
    dev = rec.device
    t_f = torch.empty(B, Nt, device=dev); pos_f = torch.empty(B, Nt, 3, device=dev); dir_f = torch.empty(B, Nt, 3, device=dev)
    grad_f = torch.empty(B, Nt, 3, device=dev) if want_grad else None
    check(_lib.load().rnerf_resample(_p(rec), W, _p(tcol), B, S, _p(t_c), _p(weights_c), Nc, _p(u), per_ray, n_fine,
                                     _p(t_f), _p(pos_f), _p(dir_f), _p(grad_f), _stream()), "rnerf_resample")
    return t_f, pos_f, dir_f, grad_f


def bbox_tail_mask(pos: torch.Tensor, lo, hi):
    """rnerf/models.py:498-503: mask = reverse-cumsum(inside bbox) > 0; returns (mask, 1 - mask) [B,Ns]."""
    pos = _chk(pos, "pos")
    B, Ns, _ = pos.shape
    m = torch.empty(B, Ns, device=pos.device); im = torch.empty(B, Ns, device=pos.device)
    check(_lib.load().rnerf_bbox_tail_mask(_p(pos), B, Ns, Dbl3(*[float(v) for v in lo]), Dbl3(*[float(v) for v in hi]),
                                           _p(m), _p(im), _stream()), "rnerf_bbox_tail_mask")
    return m, im


def encmlp_fwd_profile(packed: torch.Tensor, pos: torch.Tensor, dirs: torch.Tensor):
    """Development aid: forward + per-layer clock64 stamps of CTA 0 -> (raw, prof[2,10,4] int64)."""
    pos = _chk(pos, "pos").reshape(-1, 3); dirs = _chk(dirs, "dirs").reshape(-1, 3)
    M = pos.shape[0]
    raw = torch.empty(M, 4, device=pos.device, dtype=torch.float32)
    prof = torch.zeros(8, 10, 4, device=pos.device, dtype=torch.int64)
    check(_lib.load().rnerf_encmlp_fwd_profile(_p(packed), _p(pos), _p(dirs), M, _p(raw), _p(prof), _stream()),
          "rnerf_encmlp_fwd_profile")
    return raw, prof


# ---------------------------------------------------------------- training-mode MLP: forward with saved activations, backward
def encmlp_fwd_train(packed: torch.Tensor, pos: torch.Tensor, dirs: torch.Tensor):
    """Forward that keeps what the backward kernels need: every layer's post-activation output (bf16 [10,M,256]; the weight
    gradients' X operand), the two encodings (bf16 [2,M,64]) and the ReLU bit-masks (int32 [10,M,8]; all the dgrad chain needs
    of the activations).  Returns (raw [M,4], (layers, enc, masks))."""
    _chk(packed, "packed", torch.uint8)
    pos = _chk(pos, "pos").reshape(-1, 3); dirs = _chk(dirs, "dirs").reshape(-1, 3)
    M = pos.shape[0]
    raw = torch.empty(M, 4, device=pos.device, dtype=torch.float32)
    layers = torch.empty(10, M, 256, device=pos.device, dtype=torch.bfloat16)
    enc = torch.empty(2, M, 64, device=pos.device, dtype=torch.bfloat16)
    masks = torch.empty(10, M, 8, device=pos.device, dtype=torch.int32)
    check(_lib.load().rnerf_encmlp_fwd_train(_p(packed), _p(pos), _p(dirs), M, _p(raw), _p(layers), _p(enc), _p(masks), _stream()),
          "rnerf_encmlp_fwd_train")
    return raw, (layers, enc, masks)


def mlp_dgrad_pack(kernels) -> torch.Tensor:
    """Transposed-weight image for the dgrad chain (rebuilt whenever the weights change)."""
    lib = _lib.load()
    out = torch.empty(lib.rnerf_mlp_dgrad_packed_bytes(), device=kernels[0].device, dtype=torch.uint8)
    kp = (C.c_void_p * 12)(*[_chk(k, "kernel").data_ptr() for k in kernels])
    check(lib.rnerf_mlp_dgrad_pack(kp, _p(out), _stream()), "rnerf_mlp_dgrad_pack")
    return out


def mlp_dgrad(dgrad_packed: torch.Tensor, packed: torch.Tensor, masks: torch.Tensor, d_raw: torch.Tensor) -> torch.Tensor:
    """The dgrad chain: dZ [10, M, 256] bf16 (gradient wrt every MMA layer's pre-activation; layer 9 uses 128 columns) from
    d_raw [M, 4] and the forward's ReLU bit-masks.  Batches of >= 37 888 samples run the CTA-pair kernel
    (csrc/mlp_dgrad_pair.cu; RNERF_DGRAD_KERNEL=single forces the one-CTA kernel -- same results bit for bit)."""
    M = d_raw.shape[0]
    dz = torch.empty(10, M, 256, device=d_raw.device, dtype=torch.bfloat16)
    check(_lib.load().rnerf_mlp_dgrad(_p(dgrad_packed), _p(packed), _p(_chk(masks, "relu masks", torch.int32)),
                                      _p(_chk(d_raw.contiguous(), "d_raw")), M, _p(dz), _stream()), "rnerf_mlp_dgrad")
    return dz


def mlp_wgrad(x: torch.Tensor, x_cols: int, kx_valid: int, dz: torch.Tensor, n: int, gw: torch.Tensor,
              gb: Optional[torch.Tensor]) -> None:
    """gw[kx_valid, n] += x[:, :x_cols]^T dz[:, :n];  gb[n] += colsum(dz[:, :n]).  x: bf16 [M, ldx], dz: bf16 [M, 256]."""
    assert x.dtype == torch.bfloat16 and dz.dtype == torch.bfloat16 and x.is_contiguous() and dz.is_contiguous()
    assert gw.dtype == torch.float32 and gw.is_contiguous() and gw.shape == (kx_valid, n)
    M = x.shape[0]
    check(_lib.load().rnerf_mlp_wgrad(_p(x), x.shape[1], int(x_cols), int(kx_valid), _p(dz), int(n), M, _p(gw), _p(gb), _stream()),
          "rnerf_mlp_wgrad")


def mlp_wgrad_batched(jobs, M: int) -> None:
    """All weight-gradient GEMMs of one MLP's backward in one launch.  jobs: list of (x, x_cols, kx_valid, dz, n, gw, gb) with
    the meaning of mlp_wgrad's arguments (x / dz may be column-offset views of [M, ld] bf16 tensors)."""
    nj = len(jobs)
    xs, ldx, xc, kv, dzs, ns, gws, gbs = [], [], [], [], [], [], [], []
    for x, x_cols, kx_valid, dz, n, gw, gb in jobs:
        assert x.dtype == torch.bfloat16 and dz.dtype == torch.bfloat16 and x.is_contiguous() and dz.is_contiguous()
        assert gw.dtype == torch.float32 and gw.is_contiguous() and gw.shape == (kx_valid, n) and x.shape[0] == M and dz.shape == (M, 256)
        xs.append(x.data_ptr()); ldx.append(x.shape[1]); xc.append(int(x_cols)); kv.append(int(kx_valid))
        dzs.append(dz.data_ptr()); ns.append(int(n)); gws.append(gw.data_ptr()); gbs.append(0 if gb is None else gb.data_ptr())
    vp, ip = C.c_void_p * nj, C.c_int * nj
    check(_lib.load().rnerf_mlp_wgrad_batched(nj, vp(*xs), ip(*ldx), ip(*xc), ip(*kv), vp(*dzs), ip(*ns), int(M), vp(*gws), vp(*gbs),
                                              _stream()), "rnerf_mlp_wgrad_batched")


def mlp_input_grad_pack(k0: torch.Tensor, k5: torch.Tensor, k10: torch.Tensor) -> torch.Tensor:
    """Transposed [640][64] fp32 image of the weight rows the encodings multiply (Dense_0, Dense_5[256:], Dense_10[256:])."""
    lib = _lib.load()
    wt = torch.empty(lib.rnerf_mlp_input_grad_packed_floats(), device=k0.device, dtype=torch.float32)
    check(lib.rnerf_mlp_input_grad_pack(_p(_chk(k0, "Dense_0.kernel")), _p(_chk(k5, "Dense_5.kernel")),
                                        _p(_chk(k10, "Dense_10.kernel")), _p(wt), _stream()), "rnerf_mlp_input_grad_pack")
    return wt


def encmlp_bwd(packed, pos, dirs, saved, d_raw, params, grad_out=None, input_grads: bool = False, split: bool = False):
    """Backward of pos_enc + NerfMLP wrt the 12 Dense layers: fused tcgen05 dgrad chain (dZ of every layer), then one
    MN-major tcgen05 wgrad per layer (+ the two skinny heads on CUDA cores).  Returns [gK0, gb0, ..., gK11, gb11].
    `grad_out`: optional list of 24 fp32 tensors (same order) the kernels ACCUMULATE into -- the gradient views of a
    flat parameter arena, where every (kernel, bias) pair is contiguous; fresh zero buffers otherwise.
    `input_grads`: also return (d_pos, d_dirs), the gradients wrt the sample positions / directions ("all" stage), as
    the last element of the returned list.
    `split` (with `grad_out`): only the dgrad chain (and the input gradients) run now; returns (list as above, weight_part)
    where weight_part() launches the weight-gradient and head kernels on whatever stream is current when it is called (the
    caller orders that stream after this call and keeps it from outliving the step: train._step_body)."""
    layers, enc, masks = saved
    M = layers.shape[1]
    lib = _lib.load()
    K = [p for p in params[0::2]]
    dev = layers.device
    d_raw = _chk(d_raw.contiguous(), "d_raw")
    dz = mlp_dgrad(mlp_dgrad_pack(K), packed, masks, d_raw)
    if grad_out is not None:
        gK, gB = list(grad_out[0::2]), list(grad_out[1::2])
        for g in gK + gB:
            _chk(g, "grad_out")
        contiguous_pairs = all(gB[i].data_ptr() == gK[i].data_ptr() + 4 * gK[i].numel() for i in (8, 11))
    else:
        assert not split, "split needs grad_out (the weight part has nowhere to return its gradients)"
        contiguous_pairs = False
    inputs = None
    if input_grads:
        wt = mlp_input_grad_pack(K[0], K[5], K[10])
        pos2, dirs2 = _chk(pos.reshape(-1, 3), "pos"), _chk(dirs.reshape(-1, 3), "dirs")
        d_pos, d_dirs = torch.empty_like(pos2), torch.empty_like(dirs2)
        check(lib.rnerf_mlp_input_grad(_p(dz), M, _p(wt), _p(pos2), _p(dirs2), _p(d_pos), _p(d_dirs), _stream()),
              "rnerf_mlp_input_grad")
        inputs = (d_pos.view(pos.shape), d_dirs.view(dirs.shape))

    def weight_part():
        nonlocal gK, gB
        if not contiguous_pairs:
            heads = torch.zeros(644, device=dev, dtype=torch.float32)
            head_rgb, head_sig = heads[:387], heads[387:]
        if grad_out is None:
            gK = [torch.zeros_like(k) for k in K]
            gB = [torch.zeros_like(b) for b in params[1::2]]
            gK[11], gB[11] = head_rgb[:384].view(128, 3), head_rgb[384:387]
            gK[8], gB[8] = head_sig[:256].view(256, 1), head_sig[256:257]
        jobs = [(enc[0], 64, 63, dz[0], 256, gK[0], gB[0])]
        jobs += [(layers[l - 1], 256, 256, dz[l], 256, gK[l], gB[l]) for l in (1, 2, 3, 4, 6, 7)]
        jobs += [(layers[4], 256, 256, dz[5], 256, gK[5][:256], gB[5]), (enc[0], 64, 63, dz[5], 256, gK[5][256:], None),
                 (layers[7], 256, 256, dz[8], 256, gK[9], gB[9]), (layers[8], 256, 256, dz[9], 128, gK[10][:256], gB[10]),
                 (enc[1], 32, 27, dz[9], 128, gK[10][256:], None)]
        mlp_wgrad_batched(jobs, M)          # 13 GEMMs, one launch
        if contiguous_pairs:      # (kernel, bias) of Dense_11 / Dense_8 are adjacent in the arena: accumulate in place
            check(lib.rnerf_mlp_head_grad(_p(layers), _p(d_raw), M, _p(gK[11]), _p(gK[8]), _stream()), "rnerf_mlp_head_grad")
        else:
            check(lib.rnerf_mlp_head_grad(_p(layers), _p(d_raw), M, _p(head_rgb), _p(head_sig), _stream()), "rnerf_mlp_head_grad")
            if grad_out is not None:
                gK[11].add_(head_rgb[:384].view(128, 3)); gB[11].add_(head_rgb[384:387])
                gK[8].add_(head_sig[:256].view(256, 1)); gB[8].add_(head_sig[256:257])

    if not split:
        weight_part()
    out = []
    for i in range(12):
        out += [gK[i], gB[i]]
    if inputs is not None:
        out.append(inputs)
    return (out, weight_part) if split else out


def bkgd_mlp_bwd(w, dirs, n_rays, stride, offset, d_raw, params, gw_out=None, want_d_dirs: bool = False):
    """Backward of the background MLP wrt its 5 Dense layers (CUDA: forward recompute per 32-ray tile + chain rule,
    weight gradients accumulated with atomics).  Returns [gK0, gb0, ..., gK4, gb4] (views of `gw_out`, a flat fp32
    buffer laid out like `w` that is ACCUMULATED into, when given).  `want_d_dirs`: append d_dirs [n_rays,3], the gradient
    wrt the input directions ("all" stage)."""
    _chk(w, "w"); _chk(dirs, "dirs"); d_raw = _chk(d_raw.contiguous(), "d_raw")
    gw = torch.zeros_like(w) if gw_out is None else _chk(gw_out, "gw_out")
    assert gw.numel() == w.numel()
    ptr = C.c_void_p(dirs.data_ptr() + 4 * offset)
    d_dirs = None
    if want_d_dirs:
        d_dirs = torch.empty(n_rays, 3, device=w.device, dtype=torch.float32)
        check(_lib.load().rnerf_bkgd_mlp_bwd_dirs(_p(w), ptr, n_rays, stride, _p(d_raw), _p(gw), _p(d_dirs), _stream()),
              "rnerf_bkgd_mlp_bwd_dirs")
    else:
        check(_lib.load().rnerf_bkgd_mlp_bwd(_p(w), ptr, n_rays, stride, _p(d_raw), _p(gw), _stream()), "rnerf_bkgd_mlp_bwd")
    shapes = [(27, 128), (128, 128), (128, 128), (155, 128), (128, 3)]
    ks, off = [], 0
    for sh in shapes:
        n = sh[0] * sh[1]
        ks.append(gw[off:off + n].view(sh)); off += n
    bs = []
    for n in (128, 128, 128, 128, 3):
        bs.append(gw[off:off + n]); off += n
    out = []
    for k, b in zip(ks, bs):
        out += [k, b]
    if want_d_dirs:
        out.append(d_dirs)
    return out


# ---------------------------------------------------------------- optimiser on a flat parameter arena (a17)
HYPER_FLOATS = 10   # lr, b1, b2, eps, 1-b1^t, 1-b2^t, gscale, wd_coef, grad_max_val, grad_max_norm (include/rnerf_b200.h)


def sumsq(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[0] += sum(x^2) over a flat fp32 buffer (out: 1-element fp32, zero-initialised when not given)."""
    x = _chk(x.reshape(-1), "x")
    out = torch.zeros(1, device=x.device, dtype=torch.float32) if out is None else _chk(out, "out")
    check(_lib.load().rnerf_sumsq(_p(x), x.numel(), _p(out), _stream()), "rnerf_sumsq")
    return out


def grad_sumsq(grad: torch.Tensor, theta: Optional[torch.Tensor], hyper: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[0] += squared norm of the effective (scaled, decayed, value-clipped) gradient (train.py:176-180)."""
    grad = _chk(grad, "grad"); hyper = _chk(hyper, "hyper"); _chk(out, "out")
    assert hyper.numel() >= HYPER_FLOATS
    check(_lib.load().rnerf_grad_sumsq(_p(grad), _p(theta), grad.numel(), _p(hyper), _p(out), _stream()), "rnerf_grad_sumsq")
    return out


def adam_step(theta: torch.Tensor, grad: torch.Tensor, mu: torch.Tensor, nu: torch.Tensor, hyper: torch.Tensor,
              norm_sq: Optional[torch.Tensor] = None) -> None:
    """In-place optax.adam update of theta[:n] (n = grad.numel()) with the scalars in the device array `hyper`."""
    for nm, t in (("theta", theta), ("grad", grad), ("mu", mu), ("nu", nu), ("hyper", hyper)):
        _chk(t, nm)
    n = grad.numel()
    assert theta.numel() >= n and mu.numel() == n and nu.numel() == n and hyper.numel() >= HYPER_FLOATS
    check(_lib.load().rnerf_adam_step(_p(theta), _p(grad), _p(mu), _p(nu), n, _p(hyper), _p(norm_sq), _stream()),
          "rnerf_adam_step")


# ---------------------------------------------------------------- ray generation / image error
def generate_rays(camtoworld, height: int, width: int, focal: Optional[float] = None, cam_mat=None,
                  use_pixel_centers: bool = True, row0: int = 0, n_rows: Optional[int] = None, device="cuda",
                  want_radii: bool = True):
    """Dataset._generate_rays for one camera, on the device (rnerf/datasets.py:216-242 Blender when `focal` is given,
    :486-518 OpenCV when `cam_mat` (3x3 intrinsics) is).  Returns (origins, directions, viewdirs, radii) with shapes
    [n_rows, W, 3] / [n_rows, W, 1] for image rows [row0, row0 + n_rows)."""
    import numpy as np
    c2w = np.asarray(camtoworld, dtype=np.float64)[:3, :4]
    pose = (C.c_double * 12)(*c2w.reshape(-1).tolist())
    n_rows = height - row0 if n_rows is None else int(n_rows)
    if (focal is None) == (cam_mat is None):
        raise ValueError("give exactly one of focal (Blender camera) or cam_mat (OpenCV camera)")
    if cam_mat is not None:
        K = np.asarray(cam_mat, dtype=np.float64)
        opencv, fx, fy, cx, cy = 1, float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    else:
        opencv, fx, fy, cx, cy = 0, float(focal), float(focal), 0.0, 0.0
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.RnerfError("generate_rays: expected a CUDA device (there is no CPU implementation)")
    o = torch.empty(n_rows, width, 3, device=dev); d = torch.empty(n_rows, width, 3, device=dev)
    v = torch.empty(n_rows, width, 3, device=dev)
    r = torch.empty(n_rows, width, 1, device=dev) if want_radii else None
    with torch.cuda.device(dev):
        check(_lib.load().rnerf_generate_rays(pose, int(height), int(width), opencv, fx, fy, cx, cy, int(use_pixel_centers),
                                              int(row0), n_rows, _p(o), _p(d), _p(v), _p(r), _stream()), "rnerf_generate_rays")
    return o, d, v, r


def image_mse(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """mean((a - b)^2) as a 0-d device tensor (compute_psnr's input, rnerf/utils.py:392-401)."""
    a = _chk(a.reshape(-1), "a"); b = _chk(b.reshape(-1), "b")
    assert a.numel() == b.numel()
    out = torch.zeros(1, device=a.device, dtype=torch.float32)
    check(_lib.load().rnerf_sq_err(_p(a), _p(b), a.numel(), _p(out), _stream()), "rnerf_sq_err")
    return out[0] / a.numel()


def _env_dims(env):
    """(patch, channels) of an env map [P, P, C] (C = 3 for a whole patch; a [P/N, P, 3] shard reshaped like the reference
    does, train.py:127-128, has C = 3 N)."""
    if env is None:
        return 0, 0
    assert env.dim() == 3 and env.shape[0] == env.shape[1], f"env map must be [P, P, C], got {tuple(env.shape)}"
    return int(env.shape[0]), int(env.shape[2])


def radiance_loss_fwd(rgb, rgb_c, trb, trans, px, env, bg_weight: float, bg_smooth_weight: float, gate: float) -> torch.Tensor:
    """The radiance-stage loss terms of train.py:75-162 in one launch.  rgb / rgb_c / px: [B, 3]; trb (trans_rgb_bkgd [B, 3])
    and trans [B, 1] or None; env [P, P, C] or None.  Returns out [8] = (total, loss, loss_c, loss_bg, loss_bg_smooth, psnr,
    psnr_c, sum(mask)), layout in include/rnerf_b200.h."""
    lib = _lib.load()
    B = rgb.shape[0]
    ws = torch.zeros(lib.rnerf_radiance_loss_ws_floats(), device=rgb.device, dtype=torch.float32)
    out = torch.empty(8, device=rgb.device, dtype=torch.float32)
    for t in (rgb, rgb_c, px):
        assert _chk(t, "loss input").shape == (B, 3)
    check(lib.rnerf_radiance_loss_fwd(_p(rgb), _p(rgb_c), _p(None if trb is None else _chk(trb, "trans_rgb_bkgd")),
                                      _p(None if trb is None else _chk(trans, "trans")), _p(px), B,
                                      _p(None if env is None else _chk(env, "env")), *_env_dims(env),
                                      float(bg_weight), float(bg_smooth_weight), float(gate), _p(ws), _p(out), _stream()),
          "rnerf_radiance_loss_fwd")
    return out


def radiance_loss_bwd(rgb, rgb_c, trb, trans, px, env, bg_weight: float, bg_smooth_weight: float, gate: float, out: torch.Tensor,
                      g_total: torch.Tensor):
    """Gradients of radiance_loss_fwd's total times the device scalar g_total: (d_rgb, d_rgb_c, d_trb | None, d_env | None)."""
    lib = _lib.load()
    B = rgb.shape[0]
    d_rgb, d_rgb_c = torch.empty_like(rgb), torch.empty_like(rgb_c)
    d_trb = None if trb is None else torch.empty_like(trb)
    d_env = None if env is None else torch.empty_like(env)
    g = _chk(g_total.reshape(1).contiguous(), "g_total")
    check(lib.rnerf_radiance_loss_bwd(_p(rgb), _p(rgb_c), _p(trb), _p(trans), _p(px), B, _p(env),
                                      *_env_dims(env), float(bg_weight), float(bg_smooth_weight),
                                      float(gate), _p(out), _p(g), _p(d_rgb), _p(d_rgb_c), _p(d_trb), _p(d_env), _stream()),
          "rnerf_radiance_loss_bwd")
    return d_rgb, d_rgb_c, d_trb, d_env
