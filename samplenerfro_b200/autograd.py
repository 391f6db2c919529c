"""Autograd boundary: one torch.autograd.Function per CUDA stage.  PyTorch only records the graph between
the stages; every forward and backward body is a call into librnerf_b200.so.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops


def _mlp_param_list(p: Dict, n: int):
    out = []
    for i in range(n):
        out += [p[f"Dense_{i}"]["kernel"], p[f"Dense_{i}"]["bias"]]
    return out


def _needs_grad(p: Dict) -> bool:
    return torch.is_grad_enabled() and any(t.requires_grad for d in p.values() for t in d.values())


# ----------------------------------------------------------------------------- radiance MLP (a8 + a9)
def _sink(model, name):
    """Gradient views of a flat parameter arena (train.ParamArena) the backward kernels accumulate into directly,
    or None: the gradients are then returned to autograd as fresh tensors."""
    sinks = getattr(model, "_grad_sink", None)
    return None if sinks is None else sinks.get(name)


class _RadianceMLP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sink, fork, packed, pos, dirs, *params):
        raw, (layers, enc, masks) = ops.encmlp_fwd_train(packed, pos, dirs)
        ctx.sink = sink
        ctx.fork = fork
        ctx.save_for_backward(packed, pos, dirs, layers, enc, masks, *params)
        return raw.view(pos.shape[0], pos.shape[1], 4)

    @staticmethod
    def backward(ctx, d_raw):
        packed, pos, dirs, layers, enc, masks, *params = ctx.saved_tensors
        want_in = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]      # "all" stage: the samples depend on so3_mlp
        if ctx.fork is not None and ctx.sink is not None and not want_in:
            # Radiance stage inside a training step (train._step_body): this backward is a leaf -- it only accumulates into
            # the arena's gradient views -- so it is forked onto a side stream and joined before the gradient all-reduce.
            # The other MLP's backward, the composite and background backward kernels then run beside it: per-rank
            # batches of a multi-GPU step (512 rays) leave every one of these launches well under a wave.
            side, keep, after, defer = ctx.fork
            d_raw = d_raw.contiguous()

            def launch(ev=None):
                # ev: an event on the autograd stream after which everything this backward reads exists (deferred launch),
                # or None: order the side stream after the current one right here
                if ev is None:
                    side.wait_stream(torch.cuda.current_stream())
                else:
                    side.wait_event(ev)
                with torch.cuda.stream(side):
                    ops.encmlp_bwd(packed, pos, dirs, (layers, enc, masks), d_raw.view(-1, 4), params, grad_out=ctx.sink)
                    if after is not None:      # multi-GPU: this bucket's all-reduce starts now, under the rest of the backward
                        after()

            keep.append((side, d_raw, packed, pos, dirs, layers, enc, masks))   # alive until the join (no cross-stream reuse)
            if defer is not None:
                defer.append(launch)           # "all" stage: launched right behind the reverse sweep (_MarchAll.backward)
            else:
                launch()
            return (None, None, None, None, None) + (None,) * len(params)
        if ctx.fork is not None and ctx.sink is not None and want_in and ctx.fork[3] is not None:
            # "all" stage, coarse MLP: the sweep waits for d pos / d dirs, so the dgrad chain and the input gradients run now;
            # the weight gradients are leaves and go behind the sweep like the fine MLP's whole backward
            side, keep, after, defer = ctx.fork
            d_raw = d_raw.contiguous()
            grads, weight_part = ops.encmlp_bwd(packed, pos, dirs, (layers, enc, masks), d_raw.view(-1, 4), params,
                                                grad_out=ctx.sink, input_grads=True, split=True)
            d_pos, d_dirs = grads.pop()

            def launch_weights(ev=None):
                if ev is None:
                    side.wait_stream(torch.cuda.current_stream())
                else:
                    side.wait_event(ev)
                with torch.cuda.stream(side):
                    weight_part()
                    if after is not None:
                        after()

            keep.append((side, d_raw, packed, pos, dirs, layers, enc, masks, weight_part))
            defer.append(launch_weights)
            return (None, None, None, d_pos, d_dirs) + (None,) * len(params)
        grads = ops.encmlp_bwd(packed, pos, dirs, (layers, enc, masks), d_raw.contiguous().view(-1, 4), params,
                               grad_out=ctx.sink, input_grads=want_in)
        d_pos, d_dirs = grads.pop() if want_in else (None, None)
        if ctx.sink is not None:        # already accumulated into the arena's .grad views
            return (None, None, None, d_pos, d_dirs) + (None,) * len(params)
        return (None, None, None, d_pos, d_dirs, *grads)


def radiance_mlp(model, variables: Dict, name: str, pos: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """pos_enc + NerfMLP on [B,Ns,3] positions / per-sample directions -> raw [B,Ns,4]."""
    p = variables["params"][name]
    packed = model._packed(variables, name)
    if _needs_grad(p):
        forks = getattr(model, "_bwd_fork", None)
        return _RadianceMLP.apply(_sink(model, name), None if forks is None else forks.get(name), packed, pos, dirs,
                                  *_mlp_param_list(p, 12))
    with torch.no_grad():
        return ops.encmlp_fwd(packed, pos, dirs).view(pos.shape[0], pos.shape[1], 4)


# ----------------------------------------------------------------------------- background MLP (a10)
BKGD_TC_MIN_RAYS = 8192      # evaluations of at least this many rays run their forward on the tensor pipe (fp16 hi/lo split, fp32-grade)


def _bkgd_tc_enabled() -> bool:
    import os
    return os.environ.get("RNERF_BKGD_TC", "1") != "0"      # development aid: RNERF_BKGD_TC=0 keeps the fp32 CUDA-core kernel


class _BkgdMLP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sink, w, dirs, n_rays, stride, offset, *params):
        ctx.geom = (n_rays, stride, offset)
        ctx.sink = sink
        ctx.save_for_backward(w, dirs, *params)
        if n_rays >= BKGD_TC_MIN_RAYS and _bkgd_tc_enabled():
            # (a training step's env patch: the backward kernel recomputes the forward in fp32 anyway, so only the outputs
            # are needed here; the tensor-pipe image is rebuilt from this step's weights, two small kernels)
            return ops.bkgd_mlp_fwd_tc(ops.bkgd_tc_pack(w), dirs, n_rays, stride, offset)
        return ops.bkgd_mlp_fwd(w, dirs, n_rays, stride, offset)

    @staticmethod
    def backward(ctx, d_raw):
        w, dirs, *params = ctx.saved_tensors
        n_rays, stride, offset = ctx.geom
        want_in = ctx.needs_input_grad[2]                                # "all" stage: ray_dir_c[:, -1] depends on so3_mlp
        grads = ops.bkgd_mlp_bwd(w, dirs, n_rays, stride, offset, d_raw.contiguous(), params, gw_out=ctx.sink,
                                 want_d_dirs=want_in)
        d_dirs = None
        if want_in:
            d_dirs = torch.zeros_like(dirs)
            d_dirs.view(n_rays, -1)[:, offset:offset + 3] = grads.pop()
        if ctx.sink is not None:
            return (None, None, d_dirs) + (None,) * (3 + len(params))
        return (None, None, d_dirs, None, None, None, *grads)


def bkgd_raw(model, variables: Dict, dir_c: torch.Tensor, n_rays: int, n_coarse: int) -> torch.Tensor:
    """raw_bkgd = bkgd_mlp(pos_enc(ray_dir_c[:, -1])) (rnerf/models.py:303)."""
    p = variables["params"]["bkgd_mlp"]
    w = model._packed(variables, "bkgd_mlp")
    stride, offset = n_coarse * 3, (n_coarse - 1) * 3
    if _needs_grad(p):
        return _BkgdMLP.apply(_sink(model, "bkgd_mlp"), w, dir_c, n_rays, stride, offset, *_mlp_param_list(p, 5))
    with torch.no_grad():
        if n_rays >= BKGD_TC_MIN_RAYS and _bkgd_tc_enabled():     # a frame's worth of rays: the so3 evaluator on the tensor pipe
            return ops.bkgd_mlp_fwd_tc(model._bkgd_tc(variables), dir_c, n_rays, stride, offset)
        return ops.bkgd_mlp_fwd(w, dir_c, n_rays, stride, offset)


def bkgd_color(model, variables: Dict, viewdirs: torch.Tensor) -> torch.Tensor:
    """NerfModel.forward_envmap (rnerf/models.py:181-191): widened sigmoid of bkgd_mlp(pos_enc(dir))."""
    p = variables["params"]["bkgd_mlp"]
    w = model._packed(variables, "bkgd_mlp")
    n = viewdirs.shape[0]
    if _needs_grad(p):
        raw = _BkgdMLP.apply(_sink(model, "bkgd_mlp"), w, viewdirs, n, 3, 0, *_mlp_param_list(p, 5))
    else:
        with torch.no_grad():
            if n >= BKGD_TC_MIN_RAYS and _bkgd_tc_enabled():
                raw = ops.bkgd_mlp_fwd_tc(model._bkgd_tc(variables), viewdirs, n, 3, 0)
            else:
                raw = ops.bkgd_mlp_fwd(w, viewdirs, n, 3, 0)
    return torch.sigmoid(raw) * (1 + 2 * model.rgb_padding) - model.rgb_padding


# ----------------------------------------------------------------------------- "all"-stage march (a4-a7, trainable so3_mlp)
SO3_SORT_MAX_RAYS = 16384


class _GridTable(torch.autograd.Function):
    """(n, grad n) table of a LEARNED IoR grid (extension, no reference counterpart): forward = VoxMLP.setup's table
    (rnerf_grid_table), backward = its adjoint (rnerf_grid_table_bwd)."""

    @staticmethod
    def forward(ctx, model, grid_n):
        ctx.model = model
        # rebuilt in place in the model's own 16 B/voxel buffer (2 GiB at 512^3): no per-step allocation, graph-capturable
        ops.grid_table(grid_n, model.ndim, model.nmin, model.nmax, out=model.table)
        return model.table.view_as(model.table)

    @staticmethod
    def backward(ctx, d_table):
        m = ctx.model
        return None, ops.grid_table_bwd(d_table.contiguous(), m.ndim, m.nmin, m.nmax)


def grid_table(model, grid_n: torch.Tensor) -> torch.Tensor:
    return _GridTable.apply(model, grid_n)


def _activity_order(model, table, bricks, origins, viewdirs):
    """Permutation that groups rays by the march steps at which they need so3_mlp (first and last step with |grad n| > 1e-3,
    found by a radiance-stage march: the entry step is not affected by the rotation).  A CTA of the so3 kernels
    evaluates the MLP at the union of its rays' active steps, so a batch of random pixels costs every CTA nearly the whole
    active range; sorted, CTAs of rays that miss the object do none and the others only their own short range."""
    path = ops.march(table, model.ndim, model.nmin, model.nmax, origins, viewdirs, model.near, model.far,
                     model.num_march_steps, bricks=bricks, compact=False, t_col=False)
    act = path.rec[..., 8:11].norm(dim=-1) > 1e-3
    S = act.shape[1]
    k = torch.arange(S, device=act.device)
    first = torch.where(act, k, S).amin(dim=1)
    last = torch.where(act, k, -1).amax(dim=1)
    return torch.argsort(first * (S + 1) + last + 1)


class _MarchAll(torch.autograd.Function):
    """PathSampler + coarse selection with so3_mlp in the loop.  Differentiable outputs: pos_c, dir_c (the only way a
    loss reaches the scan: ray_dist is stop_gradient, rnerf/eikonal_utils.py:120, and so are the fine samples,
    rnerf/model_utils.py:406-411).  Backward = the reverse sweep kernel (rnerf_march_all_bwd): gradients of so3_mlp and,
    when `table` requires grad (learned IoR grid, extension), of the table.  `w` is None in the radiance stage."""

    @staticmethod
    def forward(ctx, model, sink, w, window, origins, viewdirs, jitter, compact, table, bricks, *params):
        import os
        perm = None
        so3 = (w, window) if w is not None else None
        if so3 is not None and 256 < origins.shape[0] <= SO3_SORT_MAX_RAYS and os.environ.get("RNERF_SO3_SORT", "1") != "0":
            perm = _activity_order(model, table, bricks, origins, viewdirs)
            origins, viewdirs = origins[perm].contiguous(), viewdirs[perm].contiguous()
        # the forward evaluates so3_mlp on the tensor pipe (fp16 hi/lo split operands, fp32-grade); the hi/lo weight image is
        # rebuilt from this step's weights (one small kernel)
        so3_tc = ops.so3_tc_pack(w) if so3 is not None else None
        # ... and leaves the hidden activations of every evaluation for the reverse sweep (2 KB per evaluated (ray, step))
        saved = None
        if so3 is not None and os.environ.get("RNERF_SO3_SAVE", "1") != "0":
            saved = ops.so3_saved_buffer(origins.shape[0], model.num_march_steps, origins.device)
        path = ops.march(table, model.ndim, model.nmin, model.nmax, origins, viewdirs, model.near, model.far,
                         model.num_march_steps, bricks=bricks, compact=compact, so3=so3, so3_tc=so3_tc, so3_saved=saved)
        pos_c, dir_c, t_c, _ = ops.select(path, jitter)
        ctx.model, ctx.sink, ctx.window = model, sink, window
        rec, t_col = path.rec, path.t
        if perm is not None:                  # hand everything back in the caller's ray order; the sweep keeps the sorted copy
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(perm.numel(), device=perm.device)
            pos_c, dir_c, t_c, rec, t_col = pos_c[inv], dir_c[inv], t_c[inv], rec[inv], t_col[inv]
        none = jitter.new_empty(0)
        ctx.save_for_backward(w if w is not None else none, path.rec, jitter, perm if perm is not None else none, table,
                              bricks if bricks is not None else none, saved if saved is not None else none)
        ctx.mark_non_differentiable(t_c, rec, t_col)
        return pos_c, dir_c, t_c, rec, t_col

    @staticmethod
    def backward(ctx, d_pos_c, d_dir_c, _dt, _drec, _dtcol):
        w, rec, jitter, perm, table, bricks, saved = ctx.saved_tensors
        m = ctx.model

        def z(g):
            g = torch.zeros(rec.shape[0], jitter.numel(), 3, device=rec.device) if g is None else g
            return g[perm] if perm.numel() else g

        d_table = torch.zeros_like(table) if ctx.needs_input_grad[8] else None
        so3 = (w, ctx.window) if w.numel() else None
        # training step: MLP backward work that nothing downstream waits for was deferred to here -- it is launched right
        # BEHIND the sweep (whose CTAs are latency-bound chains that free their SMs at very different times) on its side
        # streams, ordered after an event recorded BEFORE the sweep, so the two run side by side
        deferred = getattr(m, "_bwd_defer", None)
        ev = None
        if deferred:
            ev = torch.cuda.Event()
            ev.record()
        g, _, _ = ops.march_all_bwd(table, m.ndim, m.nmin, m.nmax, rec, m.near, m.far, jitter, z(d_pos_c), z(d_dir_c), so3,
                                    bricks=bricks if bricks.numel() else None, g_so3=ctx.sink if so3 is not None else None,
                                    d_table=d_table, so3_saved=saved if saved.numel() else None)
        if deferred:
            for launch in deferred:
                launch(ev)
            deferred.clear()
        n_par = len(ctx.needs_input_grad) - 10
        if ctx.sink is not None or so3 is None:
            return (None,) * 8 + (d_table, None) + (None,) * n_par
        return (None,) * 8 + (d_table, None) + tuple(ops.so3_unpack_views(g))


def march_all(model, variables: Dict, origins, viewdirs, jitter, window, compact: bool, table=None, bricks=None):
    """-> (BentPath, pos_c, dir_c, t_c) with autograd edges from pos_c / dir_c to so3_mlp ("all" stage) and to `table` when
    it requires grad.  `window`: model.so3_window(annealed_alpha) as 10 floats, or a CUDA tensor [10] read at run time."""
    table = model.table if table is None else table
    bricks = model.bricks if bricks is None and table is model.table else bricks
    if model.stage.startswith("all"):
        p = variables["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]
        w, plist, sink = model._so3_packed(variables), _mlp_param_list(p, 5), _sink(model, "so3_mlp")
    else:
        w, window, plist, sink = None, None, [], None
    pos_c, dir_c, t_c, rec, t_col = _MarchAll.apply(model, sink, w, window, origins, viewdirs, jitter, compact, table, bricks, *plist)
    return ops.BentPath(rec, t_col), pos_c, dir_c, t_c


# ----------------------------------------------------------------------------- compositing (a11 + a12)
class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, t, dirs, bkgd_raw, mask, white_bkgd, rgb_padding, sigma_bias, want_alpha=False):
        o = ops.composite_fwd(raw, t, dirs, bkgd_raw, mask, white_bkgd, rgb_padding, sigma_bias, want_weights=True,
                              want_alpha=want_alpha)
        ctx.cfg = (white_bkgd, rgb_padding, sigma_bias, bkgd_raw is not None, mask is not None)
        ctx.save_for_backward(raw, t, dirs, bkgd_raw if bkgd_raw is not None else raw.new_empty(0),
                              mask if mask is not None else raw.new_empty(0))
        # alpha feeds only the online-sparsity term, which train.py:156-159 multiplies by annealing_rate = 0:
        # its gradient contribution is exactly zero, so it is handed out as a non-differentiable output
        alpha = o["alpha"] if want_alpha else raw.new_empty(0)
        outs = (o["comp_rgb"], o["distance"], o["acc"], o["weights"], o["trans"], o["trans_rgb_bkgd"], alpha)
        ctx.mark_non_differentiable(o["distance"], o["acc"], o["weights"], alpha)
        return outs

    @staticmethod
    def backward(ctx, d_rgb, _d_dist, _d_acc, _d_w, d_trans, d_trb, _d_alpha=None):
        raw, t, dirs, bk, mask = ctx.saved_tensors
        white_bkgd, rgb_padding, sigma_bias, has_bk, has_mask = ctx.cfg
        d_raw, d_bk = ops.composite_bwd(raw, t, dirs, bk if has_bk else None, mask if has_mask else None,
                                        None if d_rgb is None else d_rgb.contiguous(),
                                        None if d_trans is None else d_trans.contiguous().view(-1),
                                        None if d_trb is None else d_trb.contiguous(),
                                        white_bkgd, rgb_padding, sigma_bias)
        return d_raw, None, None, (d_bk if has_bk else None), None, None, None, None, None


def composite(raw, t, dirs, bkgd_raw, mask, white_bkgd, rgb_padding, sigma_bias, want_weights=True, want_alpha=False):
    """activations + volumetric_rendering -> dict.  Differentiable wrt raw / bkgd_raw through comp_rgb, trans and
    trans_rgb_bkgd (the outputs train.py's loss reads); distance / acc / weights are not differentiated."""
    if torch.is_grad_enabled() and (raw.requires_grad or (bkgd_raw is not None and bkgd_raw.requires_grad)):
        rgb, dist, acc, w, trans, trb, alpha = _Composite.apply(raw, t, dirs, bkgd_raw, mask, white_bkgd, rgb_padding,
                                                                sigma_bias, bool(want_alpha))
        return {"comp_rgb": rgb, "distance": dist, "acc": acc, "weights": w, "alpha": alpha if want_alpha else None,
                "trans": trans, "trans_rgb_bkgd": trb}
    with torch.no_grad():
        return ops.composite_fwd(raw, t, dirs, bkgd_raw, mask, white_bkgd, rgb_padding, sigma_bias,
                                 want_weights=want_weights, want_alpha=want_alpha)


# ----------------------------------------------------------------------------- radiance-stage loss (train.py:75-162)
class _RadianceLoss(torch.autograd.Function):
    """loss + loss_c + bg_weight loss_bg + bg_smooth_weight loss_bg_smooth and its gradient as two kernels (csrc/loss.cu)
    instead of ~70 elementwise / reduction launches on 48 KB operands."""

    @staticmethod
    def forward(ctx, rgb, rgb_c, trb, trans, env, px, bg_weight, bg_smooth_weight, gate):
        rgb, rgb_c, px = rgb.contiguous(), rgb_c.contiguous(), px.contiguous()
        trb = None if trb is None else trb.contiguous()
        trans = None if trb is None else trans.contiguous()
        env = None if env is None else env.contiguous()
        out = ops.radiance_loss_fwd(rgb, rgb_c, trb, trans, px, env, bg_weight, bg_smooth_weight, gate)
        ctx.cfg = (bg_weight, bg_smooth_weight, gate, trb is not None, env is not None)
        e = rgb.new_empty(0)
        ctx.save_for_backward(rgb, rgb_c, trb if trb is not None else e, trans if trans is not None else e,
                              env if env is not None else e, px, out)
        total, stats = out[0], out[1:]
        ctx.mark_non_differentiable(stats)
        return total, stats

    @staticmethod
    def backward(ctx, g_total, _g_stats):
        rgb, rgb_c, trb, trans, env, px, out = ctx.saved_tensors
        bg_weight, bg_smooth_weight, gate, has_bg, has_env = ctx.cfg
        d_rgb, d_rgb_c, d_trb, d_env = ops.radiance_loss_bwd(rgb, rgb_c, trb if has_bg else None, trans if has_bg else None, px,
                                                             env if has_env else None, bg_weight, bg_smooth_weight, gate, out,
                                                             g_total.to(torch.float32))
        return d_rgb, d_rgb_c, d_trb, None, d_env, None, None, None, None


def radiance_loss(rgb, rgb_c, trb, trans, env, px, bg_weight: float, bg_smooth_weight: float, gate: float):
    """Returns (total, stats) with stats = (loss, loss_c, loss_bg, loss_bg_smooth, psnr, psnr_c, sum(mask)) detached;
    trb / trans = None drops the background term, env = None the smoothness term."""
    return _RadianceLoss.apply(rgb, rgb_c, trb, trans, env, px, float(bg_weight), float(bg_smooth_weight), float(gate))
