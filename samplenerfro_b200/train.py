"""Host-side mirror of train.py's optimisation step for the radiance and "all" stages (train.py:58-183, 290-310).

`train_step(model, rng, state, batch) -> (new_state, stats, rng)` keeps the reference's signature.  The loss is
train.py:75-162 with annealing_rate hard-coded to 0 (train.py:156); gradients are averaged across ranks
(`jax.lax.pmean(grads, "batch")`, train.py:166) with torch.distributed all-reduces bucketed per MLP, then Adam
(optax.adam defaults) with `learning_rate_decay` is applied identically on every rank.  In the "all" stage so3_mlp is a
fourth trainable bucket: its gradient comes from the reverse sweep of the eikonal scan (`autograd._MarchAll`).
"""
from __future__ import annotations

import dataclasses
import os
import math
from typing import Any, Dict, List, Optional

import torch

from . import utils

GRAD_BUCKETS = ("fine_mlp", "coarse_mlp", "bkgd_mlp")   # reverse order of backward completion
ALL_STAGE_BUCKETS = GRAD_BUCKETS + ("path_sampler",)    # train.py:302-310: the "all" stage also trains so3_mlp
IOR_STAGE_BUCKETS = ("path_sampler",)                   # train.py:295-301: the "ior" stage freezes the three radiance MLPs


def tree_leaves(tree) -> List[torch.Tensor]:
    out: List[torch.Tensor] = []
    if isinstance(tree, dict):
        for k in tree:
            out += tree_leaves(tree[k])
    else:
        out.append(tree)
    return out


# development aid: RNERF_FORK_BACKWARD=0 keeps the whole backward on one stream (A/B runs)
_FORK_BACKWARD = os.environ.get("RNERF_FORK_BACKWARD", "1") != "0"


class ParamArena:
    """Every leaf of the variables tree as a view of ONE flat fp32 buffer, trainable buckets first
    (fine_mlp | coarse_mlp | bkgd_mlp | frozen rest), with a same-layout gradient buffer whose views are pre-installed
    as the leaves' `.grad`.  One memset zeroes all gradients, one NCCL call per bucket reduces them in place (no
    flatten/unflatten copies), one kernel applies Adam, one kernel gives the weight_l2 statistic (padding stays zero).
    Inside a bucket the leaves follow Flax order (Dense_i kernel, bias), except bkgd_mlp which is laid out (kernels..., biases...): that is
    the background kernels' weight image, so the bucket itself is passed to them and no packing step exists."""

    def __init__(self, variables: Dict, buckets=GRAD_BUCKETS):
        params = variables["params"]
        self.buckets = tuple(buckets)

        def walk(d, out):
            for k in d:
                if isinstance(d[k], dict):
                    walk(d[k], out)
                else:
                    out.append((d, k))
            return out

        groups = []                                   # (bucket name, [(container dict, key), ...])
        flat_images = {"bkgd_mlp": ("bkgd_mlp", lambda: params["bkgd_mlp"]),
                       "path_sampler": ("so3_mlp", lambda: params["path_sampler"]["scan"]["idx_model"]["so3_mlp"])}
        for name in self.buckets:
            if name in flat_images:       # laid out (kernels..., biases...): the kernels' own weight image
                mlp = flat_images[name][1]()
                layers = [mlp[f"Dense_{i}"] for i in range(len(mlp))]
                groups.append((name, [(lay, leaf) for leaf in ("kernel", "bias") for lay in layers]))
            else:
                groups.append((name, walk(params[name], [])))
        frozen = []
        for name in params:
            if name not in self.buckets:
                walk(params[name], frozen)
        groups.append(("frozen", frozen))
        dev = groups[0][1][0][0][groups[0][1][0][1]].device
        self.numel = sum(d[k].numel() for _, ents in groups for d, k in ents)   # true leaf count (weight_l2 denominator)
        self.bucket_range: Dict[str, tuple] = {}
        off, layout = 0, []
        for name, ents in groups:
            off = (off + 3) // 4 * 4                                            # 16-byte aligned bucket starts
            if name == "frozen":
                self.n_train = off
            lo = off
            for d, k in ents:
                off = (off + 3) // 4 * 4          # 16-byte aligned leaves: the wgrad kernel reduces with red.v4.f32
                layout.append((d, k, off, name))
                off += d[k].numel()
            self.bucket_range[name] = (lo, off)
        total = (off + 3) // 4 * 4
        self.theta = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(self.n_train, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for d, k, o, name in layout:
                t = d[k]
                view = self.theta[o:o + t.numel()].view(t.shape)
                view.copy_(t)
                view = view.detach()
                if name != "frozen":
                    view.requires_grad_(True)
                    view.grad = self.grad[o:o + t.numel()].view(t.shape)
                d[k] = view
        self.sinks: Dict[str, Any] = {}
        for name in ("fine_mlp", "coarse_mlp"):
            p = params[name]
            self.sinks[name] = [p[f"Dense_{i}"][leaf].grad for i in range(len(p)) for leaf in ("kernel", "bias")]
        self.theta_flat = {}
        for name in self.buckets:
            if name in flat_images:
                lo, hi = self.bucket_range[name]
                assert hi - lo == sum(d[k].numel() for d, k in groups[self.buckets.index(name)][1]), f"{name} bucket must be dense"
                self.sinks[flat_images[name][0]] = self.grad[lo:hi]
                self.theta_flat[flat_images[name][0]] = self.theta[lo:hi]

    def zero_grad(self) -> None:
        self.grad.zero_()

    def bucket_grads(self) -> List[torch.Tensor]:
        return [self.grad[self.bucket_range[n][0]:self.bucket_range[n][1]] for n in self.buckets]

    def allreduce_bucket(self, name: str, group=None):
        """Start the in-place all-reduce(sum) of one bucket's gradient view on the CURRENT stream; returns the work handle."""
        import torch.distributed as dist
        lo, hi = self.bucket_range[name]
        return dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True)

    def allreduce_mean(self, world_size: int, group=None, started=()) -> None:
        """jax.lax.pmean(grads, "batch") (train.py:166): one in-place all-reduce per bucket view, then one scale.
        `started`: (bucket name, work) pairs whose all-reduce was already issued (train._step_body starts a radiance MLP's
        bucket on the side stream of its backward, so the transfer runs under the rest of the backward)."""
        if world_size <= 1:
            return
        done = {n for n, _ in started}
        works = [w for _, w in started] + [self.allreduce_bucket(n, group) for n in self.buckets if n not in done]
        for w in works:
            w.wait()
        self.grad.mul_(1.0 / world_size)

    def weight_l2(self) -> torch.Tensor:
        """mean(theta^2) over every leaf (train.py:146-150) -- one kernel over the arena (padding is zero)."""
        from . import ops
        if self.theta.is_cuda:
            return ops.sumsq(self.theta)[0] / self.numel
        return (self.theta ** 2).sum() / self.numel


class _PinnedRing:
    """Host -> device staging of small per-step values without a stream synchronisation: a few pinned slots, each
    guarded by an event so a slot is not rewritten while its copy may still be in flight."""

    def __init__(self, shape, dtype, device, slots: int = 4):
        self.cuda = torch.device(device).type == "cuda"
        self.bufs = [torch.zeros(shape, dtype=dtype) for _ in range(slots)]
        if self.cuda:
            self.bufs = [b.pin_memory() for b in self.bufs]
        self.events: List[Any] = [None] * slots
        self.i = 0

    def push(self, src: torch.Tensor, dst: torch.Tensor) -> None:
        if not self.cuda:
            dst.copy_(src)
            return
        i = self.i
        self.i = (i + 1) % len(self.bufs)
        if self.events[i] is not None:
            self.events[i].synchronize()
        self.bufs[i].copy_(src)
        dst.copy_(self.bufs[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev


class ArenaAdam:
    """optax.adam(lr) with the reference's clipping (train.py:169-181) as one kernel over the arena; the per-step scalars
    are staged in pinned memory and copied to a device array so a captured graph of the step can be replayed."""

    def __init__(self, arena: ParamArena, args):
        from . import ops
        self.arena, self.b1, self.b2, self.eps = arena, 0.9, 0.999, 1e-8
        dev = arena.theta.device
        self.mu = torch.zeros_like(arena.grad)
        self.nu = torch.zeros_like(arena.grad)
        self.hyper = torch.zeros(ops.HYPER_FLOATS, device=dev, dtype=torch.float32)
        self.ring = _PinnedRing((ops.HYPER_FLOATS,), torch.float32, dev)
        self.norm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.count = 0                               # optax's own step counter (bias correction)
        self.wd_coef = 2.0 * float(args.weight_decay_mult) / arena.numel
        self.grad_max_val, self.grad_max_norm = float(args.grad_max_val), float(args.grad_max_norm)
        self._zero_frozen = None

    def stage_hyper(self, lr: float) -> None:
        """Host -> device copy of this step's scalars (outside any captured graph)."""
        t = self.count + 1
        self.ring.push(torch.tensor([lr, self.b1, self.b2, self.eps, 1.0 - self.b1 ** t, 1.0 - self.b2 ** t, 1.0,
                                     self.wd_coef, self.grad_max_val, self.grad_max_norm], dtype=torch.float32), self.hyper)

    def apply(self) -> None:
        """The device part of the update (capturable): optional global-norm pass, then the fused Adam kernel."""
        from . import ops
        norm = None
        if self.grad_max_norm > 0:
            self.norm_sq.zero_()
            norm = ops.grad_sumsq(self.arena.grad, self.arena.theta, self.hyper, self.norm_sq)
            # train.py:176-180 takes the norm over the WHOLE grads tree, before optax.set_to_zero() discards the frozen
            # subtrees: those leaves still carry the (value-clipped) weight-decay gradient 2 wd / numel * theta
            n_frozen = self.arena.theta.numel() - self.arena.n_train
            if self.wd_coef != 0.0 and n_frozen > 0:
                if self._zero_frozen is None:
                    self._zero_frozen = torch.zeros(n_frozen, device=self.arena.theta.device, dtype=torch.float32)
                ops.grad_sumsq(self._zero_frozen, self.arena.theta[self.arena.n_train:], self.hyper, self.norm_sq)
        ops.adam_step(self.arena.theta, self.arena.grad, self.mu, self.nu, self.hyper, norm)

    def step(self, lr: float) -> None:
        self.stage_hyper(lr)
        self.apply()
        self.count += 1


class GridAdam:
    """Extension (no reference counterpart): optax.adam on a learned IoR grid, `model.grid_n` [G^3], with the MLPs' learning
    rate schedule, no weight decay and no clipping; the same fused kernel as ArenaAdam, its own moments and scalars."""

    def __init__(self, model, arena_opt: "ArenaAdam"):
        from . import ops
        self.model, self.ref = model, arena_opt
        g = model.grid_n
        g.grad = torch.zeros_like(g)
        self.mu, self.nu = torch.zeros_like(g), torch.zeros_like(g)
        self.hyper = torch.zeros(ops.HYPER_FLOATS, device=g.device, dtype=torch.float32)
        self.ring = _PinnedRing((ops.HYPER_FLOATS,), torch.float32, g.device)

    def stage_hyper(self, lr: float) -> None:
        t = self.ref.count + 1
        self.ring.push(torch.tensor([lr, self.ref.b1, self.ref.b2, self.ref.eps, 1.0 - self.ref.b1 ** t, 1.0 - self.ref.b2 ** t,
                                     1.0, 0.0, 0.0, 0.0], dtype=torch.float32), self.hyper)

    def zero_grad(self) -> None:
        self.model.grid_n.grad.zero_()

    def allreduce_mean(self, world_size: int, group=None) -> None:
        if world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.model.grid_n.grad, op=dist.ReduceOp.SUM, group=group)
            self.model.grid_n.grad.mul_(1.0 / world_size)

    def apply(self) -> None:
        from . import ops
        g = self.model.grid_n
        ops.adam_step(g.detach(), g.grad, self.mu, self.nu, self.hyper, None)


@dataclasses.dataclass
class TrainState:
    """flax.training.train_state.TrainState stand-in: step, params (the variables tree), optimiser state."""
    step: int
    params: Dict
    opt: Any = None
    arena: Optional[ParamArena] = None
    graphs: Dict = dataclasses.field(default_factory=dict)
    grid_opt: Optional[GridAdam] = None

    def replayed_kernel_launches(self) -> int:
        """Kernels of this library launched through graph replays so far (the eager ones are in _lib.launch_count())."""
        return sum(g.kernels_per_replay * g.replays for g in self.graphs.values() if isinstance(g, _GraphedStep))

    @staticmethod
    def create(variables: Dict, args, model=None) -> "TrainState":
        """Re-homes the variables into a ParamArena (the tree keeps its names; leaves become views) and attaches the
        fused Adam.  Radiance stage: path_sampler gets optax.set_to_zero (T7) -> it sits in the frozen tail; "all" stage
        (train.py:302-310): so3_mlp is a fourth trainable bucket, laid out as the march kernels' weight image."""
        stage = str(getattr(args, "stage", "radiance"))
        arena = ParamArena(variables, ALL_STAGE_BUCKETS if stage.startswith("all") else
                           (IOR_STAGE_BUCKETS if stage.startswith("ior") else GRAD_BUCKETS))
        opt = ArenaAdam(arena, args)
        grid_opt = GridAdam(model, opt) if model is not None and getattr(model, "grid_n", None) is not None else None
        return TrainState(step=0, params=variables, opt=opt, arena=arena, grid_opt=grid_opt)


def _ior_stage_loss(model, variables, batch, args, annealed_alpha, arena):
    """train.py:131-143, the "ior" stage as the reference has it: no rendering; loss_nrm = normal_loss, which
    compute_normal_loss_and_smooth returns as the constant 0.0 (rnerf/eikonal_utils.py:98), and every other term is 0 -- so
    the only thing that reaches the parameters is the weight-decay term.  The smoothness statistic is still evaluated when the
    batch carries the Grid points, like the reference does."""
    dev = model.device
    zero = torch.zeros((), device=dev)
    if batch.get("pts") is not None:
        model.apply(variables, batch["pts"], batch["grads"], annealed_alpha, method=model.wrapper_compute_normal_loss_and_smooth)
    if arena is not None:
        with torch.no_grad():
            weight_l2 = arena.weight_l2()
    else:
        leaves = tree_leaves(variables)
        weight_l2 = sum((z ** 2).sum() for z in leaves) / sum(z.numel() for z in leaves)
    total = args.weight_decay_mult * weight_l2
    stats = {"loss": zero, "psnr": zero, "loss_c": zero, "psnr_c": zero, "weight_l2": weight_l2.detach(), "loss_bg": zero,
             "loss_bg_smooth": zero, "loss_sp": zero, "loss_nrm": zero, "annealing_rate": annealed_alpha}
    return total, stats


def loss_fn(model, variables, batch, args, key_0, key_1, jitter=None, u=None, arena: Optional[ParamArena] = None):
    """train.py:75-162, radiance stage.  Returns (total, stats).  With `arena`, the weight_l2 term is a statistic only:
    its closed-form gradient is applied by the optimiser kernel (ArenaAdam)."""
    annealed_alpha = float(batch["annealed_alpha"])
    if str(getattr(args, "stage", "radiance")).startswith("ior"):
        return _ior_stage_loss(model, variables, batch, args, annealed_alpha, arena)
    rays = batch["rays"]
    env, env_stream = None, getattr(model, "_env_stream", None)
    if args.bg_smooth_weight > 0:
        vd = batch["env_rays"].viewdirs
        ps = vd.shape[0]
        if vd.shape[0] * vd.shape[1] > 4096:
            # a whole 128 x 128 patch fills the GPU on its own: beside the march and the MLP backwards it only delays their
            # persistent CTAs (measured 6.35 -> 6.59 ms at 4096 rays); the 16-row shard of an 8-GPU step gains 0.11 of 1.73 ms
            env_stream = None
        if env_stream is not None:
            # inside a training step: the env patch depends on nothing the rays produce, so its forward runs on a side stream
            # under the (latency-bound) march -- and autograd runs its backward on that stream too, beside the MLP backwards
            model._env_forked = True          # train._step_body joins the stream after backward
            env_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(env_stream):
                env = model.apply(variables, vd.reshape(-1, 3), method=model.forward_envmap).reshape(ps, ps, -1)
    ret, loss_sp = model.apply(variables, key_0, key_1, rays, args.randomized, annealed_alpha, jitter=jitter, u=u,
                               so3_window=batch.get("so3_window"))
    rgb, _d, _a, trans, trans_rgb_bkgd = ret[-1]
    px = batch["pixels"][..., :3]
    gate = 1.0 if annealed_alpha > 0 else 0.0
    rgb_c = ret[0][0]
    if args.bg_smooth_weight > 0:
        if env_stream is not None:
            torch.cuda.current_stream().wait_stream(env_stream)
        else:
            env = model.apply(variables, vd.reshape(-1, 3), method=model.forward_envmap).reshape(ps, ps, -1)
    # train.py:86-118: the image losses, the background term and the env-map smoothness term, fused with their gradient
    # (csrc/loss.cu): loss = mse(rgb), loss_c = mse(rgb_c), loss_bg = gate sum(mask |trb - px|) / (sum(mask) + 1) with
    # mask = trans > 0.5, loss_bg_smooth = gate mean(0.5 dv^2 + 0.5 dh^2) over the env patch
    from . import autograd as ag
    has_bg = args.bg_weight > 0
    image_total, ls = ag.radiance_loss(rgb, rgb_c, trans_rgb_bkgd if has_bg else None, trans if has_bg else None, env, px,
                                       args.bg_weight, args.bg_smooth_weight, gate)
    loss, loss_c, loss_bg, loss_bg_smooth, psnr, psnr_c = ls[0], ls[1], ls[2], ls[3], ls[4], ls[5]
    loss_nrm = torch.zeros((), device=rgb.device)
    annealing_rate = 0.0                      # train.py:156 (the annealed expression is commented out): loss_nrm and loss_sp are multiplied by it
    if (str(getattr(args, "stage", "radiance")).startswith("all") and batch.get("pts") is not None
            and (args.normal_loss_weight + args.normal_smooth_weight) > 0):
        # train.py:120-124: evaluated like the reference does, although annealing_rate = 0 removes it from loss and stats
        nl, ns = model.apply(variables, batch["pts"], batch["grads"], annealed_alpha, method=model.wrapper_compute_normal_loss_and_smooth,
                             so3_window=batch.get("so3_window"))
        loss_nrm = annealing_rate * (args.normal_loss_weight * nl + args.normal_smooth_weight * ns)
    if arena is not None:
        with torch.no_grad():
            weight_l2 = arena.weight_l2()
    else:
        leaves = tree_leaves(variables)
        weight_l2 = sum((z ** 2).sum() for z in leaves) / sum(z.numel() for z in leaves)
    total = image_total + args.weight_decay_mult * weight_l2
    stats = {"loss": loss, "psnr": psnr, "loss_c": loss_c, "psnr_c": psnr_c, "weight_l2": weight_l2.detach(),
             "loss_bg": args.bg_weight * loss_bg, "loss_bg_smooth": loss_bg_smooth,
             "loss_sp": torch.zeros((), device=rgb.device), "loss_nrm": loss_nrm,
             "annealing_rate": annealed_alpha}
    return total, stats


def allreduce_mean_grads(variables: Dict, world_size: int, group=None) -> None:
    """jax.lax.pmean(grads, "batch") (train.py:166): one flattened all-reduce per MLP bucket (5.26 MB total),
    issued in the order backward finishes them, then scaled by 1/N."""
    if world_size <= 1:
        return
    import torch.distributed as dist
    works = []
    for name in GRAD_BUCKETS:
        leaves = [p for p in tree_leaves(variables["params"][name]) if p.grad is not None]
        if not leaves:
            continue
        flat = torch.cat([p.grad.reshape(-1) for p in leaves])
        works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, leaves))
    for work, flat, leaves in works:
        work.wait()
        flat.div_(world_size)
        off = 0
        for p in leaves:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n


def _step_body(model, state: TrainState, batch: Dict, args, key_0, key_1, world_size, group, jitter, u):
    """Device work of one optimisation step: zero grads, loss forward/backward (gradients land in the arena),
    gradient + stats all-reduce, fused Adam.  No host synchronisation: capturable in a CUDA graph."""
    arena = state.arena
    arena.zero_grad()
    if state.grid_opt is not None:
        state.grid_opt.zero_grad()
    model._grad_sink, model._theta_flat = arena.sinks, arena.theta_flat
    keep: list = []
    started: list = []          # (bucket, work) of all-reduces issued on the side streams
    if arena.theta.is_cuda and _FORK_BACKWARD:
        # the radiance MLPs' backwards (leaves: they only accumulate into arena.grad) run on side streams next to everything
        # autograd executes after them; joined below, before the gradients are reduced (autograd._RadianceMLP.backward)
        if getattr(state, "_bwd_streams", None) is None:
            state._bwd_streams = [torch.cuda.Stream(device=arena.theta.device) for _ in range(2)]
        names = os.environ.get("RNERF_FORK_BACKWARD_MLPS", "fine_mlp,coarse_mlp").split(",")
        early = world_size > 1 and os.environ.get("RNERF_EARLY_ALLREDUCE", "1") != "0"

        def after(name):        # runs on the bucket's side stream right after its backward kernels were issued
            return (lambda: started.append((name, arena.allreduce_bucket(name, group)))) if early else None

        # "all" stage: the reverse sweep of the scan is a 14 ms chain of latency-bound CTAs; MLP backward work that nothing
        # waits for is deferred and launched behind it (autograd._MarchAll.backward), instead of ahead of it
        defer = [] if (str(getattr(args, "stage", "")).startswith("all") and "path_sampler" in arena.buckets
                       and os.environ.get("RNERF_DEFER_UNDER_SWEEP", "1") != "0") else None
        model._bwd_defer = defer
        model._bwd_fork = {n: (st, keep, after(n), defer) for n, st in zip(names, state._bwd_streams) if n in arena.buckets}
        if os.environ.get("RNERF_FORK_ENV", "1") != "0":
            if getattr(state, "_env_stream", None) is None:
                state._env_stream = torch.cuda.Stream(device=arena.theta.device)
            model._env_stream = state._env_stream
    try:
        total, stats = loss_fn(model, state.params, batch, args, key_0, key_1, jitter=jitter, u=u, arena=arena)
        if total.requires_grad:       # ("ior" stage with the arena: only the closed-form weight-decay gradient exists)
            total.backward()
    finally:
        model._grad_sink = None
        model._bwd_fork = None
        model._env_stream = None
        leftover = getattr(model, "_bwd_defer", None)
        model._bwd_defer = None
        if leftover:                    # no sweep ran (nothing reached the sampler): launch what was deferred now
            for launch in leftover:
                launch(None)
            leftover.clear()
        if keep:
            for st in {id(k[0]): k[0] for k in keep}.values():      # only the streams a backward was actually forked onto
                torch.cuda.current_stream().wait_stream(st)
            keep.clear()
        if getattr(model, "_env_forked", False):
            # the env patch's backward accumulates into arena.grad on its own stream and hands autograd no leaf gradient,
            # so the engine has nothing to synchronise on: join explicitly before the gradients are read
            torch.cuda.current_stream().wait_stream(state._env_stream)
            model._env_forked = False
    arena.allreduce_mean(world_size, group, started=started)
    if state.grid_opt is not None:
        state.grid_opt.allreduce_mean(world_size, group)
    if world_size > 1:
        import torch.distributed as dist
        keys = [k for k, v in stats.items() if torch.is_tensor(v)]
        packed = torch.stack([stats[k].float() for k in keys])
        dist.all_reduce(packed, group=group)                      # pmean(stats), train.py:167
        for k, v in zip(keys, packed / world_size):
            stats[k] = v
    state.opt.apply()
    if state.grid_opt is not None:
        state.grid_opt.apply()
    return stats


class _GraphedStep:
    """The whole training step captured once in a CUDA graph and replayed: the step is ~100 kernels of a few
    microseconds to a few hundred each, so eager launches from Python (about 15 ms of host time at 4096 rays) would
    bound it.  Static input buffers are refreshed before every replay; per-step scalars go through ArenaAdam.hyper."""

    def __init__(self, model, state, batch, args, world_size, group):
        dev = state.arena.theta.device
        rays = batch["rays"]
        B = rays.origins.shape[0]
        self.rays = utils.Rays(*[torch.empty_like(r, device=dev).contiguous() for r in rays])
        self.pixels = torch.empty_like(batch["pixels"], device=dev)
        self.env = None
        if "env_rays" in batch and batch["env_rays"] is not None:
            self.env = utils.Rays(*[torch.empty_like(r, device=dev).contiguous() for r in batch["env_rays"]])
        self.jitter = torch.zeros(model.num_coarse_samples, device=dev, dtype=torch.int32)
        self.jitter_ring = _PinnedRing((model.num_coarse_samples,), torch.int32, dev)
        self.u = (torch.zeros(B, model.num_fine_samples, device=dev) if args.randomized
                  else model.draw_u(None, B, False))
        self.static_batch = {"rays": self.rays, "pixels": self.pixels, "env_rays": self.env,
                             "annealed_alpha": float(batch["annealed_alpha"])}
        # "all" stage: the so3 positional-encoding window follows annealed_alpha (train.py:350-351), which changes every
        # step -- the captured kernels read it from this device buffer, refreshed before each replay (not baked in by value)
        self.window = None
        if str(getattr(args, "stage", "radiance")).startswith("all"):
            self.window = torch.zeros(10, device=dev, dtype=torch.float32)
            self.window_ring = _PinnedRing((10,), torch.float32, dev)
            self.static_batch["so3_window"] = self.window
        self.load(model, batch, 0, 0, args)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.stats = _step_body(model, state, self.static_batch, args, None, None, world_size, group, self.jitter, self.u)
        self.kernels_per_replay = _lib.launch_count() - l0     # this library's kernel nodes in the captured graph
        self.replays = 0
        model._pack_cache.clear()

    def load(self, model, batch, key_0, key_1, args):
        for dst, src in zip(self.rays, batch["rays"]):
            dst.copy_(src, non_blocking=True)
        self.pixels.copy_(batch["pixels"], non_blocking=True)
        if self.env is not None:
            for dst, src in zip(self.env, batch["env_rays"]):
                dst.copy_(src, non_blocking=True)
        self.jitter_ring.push(model.draw_jitter(key_0, host=True), self.jitter)
        if self.window is not None:
            self.window_ring.push(torch.tensor(model.so3_window(float(batch["annealed_alpha"])), dtype=torch.float32), self.window)
        if args.randomized:
            self.u.copy_(model.draw_u(key_1, self.u.shape[0], True))

    def replay(self):
        self.graph.replay()
        self.replays += 1
        return dict(self.stats)


def _graph_key(batch, args, world_size):
    env = batch.get("env_rays")
    return (tuple(batch["rays"].origins.shape), tuple(batch["pixels"].shape),
            None if env is None else tuple(env.viewdirs.shape), float(batch["annealed_alpha"]) > 0,
            bool(args.randomized), int(world_size))


def train_step(model, rng, state: TrainState, batch: Dict, args=None, world_size: int = 1, group=None,
               jitter=None, u=None, use_graph: Optional[bool] = None, graph_after: int = 2):
    """One optimisation step (train.py:58-183).  After `graph_after` eager steps with a given batch shape the step is
    captured in a CUDA graph and replayed from then on (`use_graph=False` keeps it eager; explicit `jitter`/`u`
    also do).  Returned stats are device tensors: reading them is the caller's synchronisation point."""
    args = args if args is not None else batch["args"]
    key_0, key_1 = utils._split_key(rng)
    lr = utils.learning_rate_decay(state.step, args.lr_init, args.lr_final, args.max_steps, args.lr_delay_steps,
                                   args.lr_delay_mult)
    on_cuda = state.arena.theta.is_cuda
    if use_graph is None:     # (the "ior" stage has no rays and a handful of kernels: nothing to capture)
        use_graph = on_cuda and jitter is None and u is None and not str(getattr(args, "stage", "")).startswith("ior")
    state.opt.stage_hyper(lr)
    if state.grid_opt is not None:
        state.grid_opt.stage_hyper(lr)
    stats = None
    if use_graph:
        key = _graph_key(batch, args, world_size)
        slot = state.graphs.get(key)
        if isinstance(slot, _GraphedStep):
            slot.load(model, batch, key_0, key_1, args)
            stats = slot.replay()
        elif isinstance(slot, int) and slot >= graph_after:
            slot = _GraphedStep(model, state, batch, args, world_size, group)
            state.graphs[key] = slot
            slot.load(model, batch, key_0, key_1, args)
            stats = slot.replay()
        else:
            state.graphs[key] = (slot or 0) + 1
    if stats is None:
        stats = _step_body(model, state, batch, args, key_0, key_1, world_size, group, jitter, u)
    model._pack_cache.clear()          # the weights changed under the packed images
    state.opt.count += 1
    state.step += 1
    stats["lr"] = lr
    stats["annealing_rate"] = float(batch["annealed_alpha"])
    return state, stats, (int(rng) + 1 if rng is not None else 1)


def shutdown_distributed(state: Optional[TrainState] = None) -> None:
    """Tear down in the order NCCL needs: captured graphs that contain all-reduce nodes keep the communicator busy, so
    they are released (and the device drained) BEFORE the process group is destroyed; destroying the group first
    blocks forever (seen on 2 x B200, NCCL 2.28)."""
    import gc
    import torch.distributed as dist
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if state is not None:
        state.graphs.clear()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
