"""Host-side mirror of train.py's optimisation step for the radiance stage (train.py:58-183).

`train_step(model, rng, state, batch) -> (new_state, stats, rng)` keeps the reference's signature.  The loss is
train.py:75-162 with annealing_rate hard-coded to 0 (train.py:156); gradients are averaged across ranks
(`jax.lax.pmean(grads, "batch")`, train.py:166) with torch.distributed all-reduces bucketed per MLP, then Adam
(optax.adam defaults) with `learning_rate_decay` is applied identically on every rank.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Any, Dict, List, Optional

import torch

from . import utils

GRAD_BUCKETS = ("fine_mlp", "coarse_mlp", "bkgd_mlp")   # reverse order of backward completion


def tree_leaves(tree) -> List[torch.Tensor]:
    out: List[torch.Tensor] = []
    if isinstance(tree, dict):
        for k in tree:
            out += tree_leaves(tree[k])
    else:
        out.append(tree)
    return out


@dataclasses.dataclass
class TrainState:
    """flax.training.train_state.TrainState stand-in: step, params (the variables tree), optimiser state."""
    step: int
    params: Dict
    opt: Any = None

    @staticmethod
    def create(variables: Dict, args) -> "TrainState":
        trainable = []
        for name in GRAD_BUCKETS:                       # radiance stage: path_sampler gets optax.set_to_zero (T7)
            trainable += tree_leaves(variables["params"][name])
        for p in trainable:
            p.requires_grad_(True)
        opt = torch.optim.Adam(trainable, lr=args.lr_init, betas=(0.9, 0.999), eps=1e-8)   # optax.adam defaults
        return TrainState(step=0, params=variables, opt=opt)


def loss_fn(model, variables, batch, args, key_0, key_1, jitter=None, u=None):
    """train.py:75-162, radiance stage.  Returns (total, stats)."""
    annealed_alpha = float(batch["annealed_alpha"])
    rays = batch["rays"]
    ret, loss_sp = model.apply(variables, key_0, key_1, rays, args.randomized, annealed_alpha, jitter=jitter, u=u)
    rgb, _d, _a, trans, trans_rgb_bkgd = ret[-1]
    px = batch["pixels"][..., :3]
    loss = ((rgb - px) ** 2).mean()
    gate = 1.0 if annealed_alpha > 0 else 0.0
    if args.bg_weight > 0:
        mask_bg = (trans > 0.5).float()
        loss_bg = gate * (mask_bg * torch.abs(trans_rgb_bkgd - px)).sum() / (mask_bg.sum() + 1)
    else:
        loss_bg = torch.zeros((), device=rgb.device)
    rgb_c = ret[0][0]
    loss_c = ((rgb_c - px) ** 2).mean()
    if args.bg_smooth_weight > 0:
        vd = batch["env_rays"].viewdirs
        ps = vd.shape[0]
        env = model.apply(variables, vd.reshape(-1, 3), method=model.forward_envmap).reshape(ps, ps, -1)
        loss_bg_smooth = gate * torch.mean(0.5 * ((env[1:, :] - env[:-1, :]) ** 2).reshape(-1)
                                           + 0.5 * ((env[:, 1:] - env[:, :-1]) ** 2).reshape(-1))
    else:
        loss_bg_smooth = torch.zeros((), device=rgb.device)
    leaves = tree_leaves(variables)
    weight_l2 = sum((z ** 2).sum() for z in leaves) / sum(z.numel() for z in leaves)
    total = (loss + loss_c + args.bg_weight * loss_bg + args.bg_smooth_weight * loss_bg_smooth
             + args.weight_decay_mult * weight_l2)
    stats = {"loss": loss.detach(), "psnr": utils.compute_psnr(loss.detach()), "loss_c": loss_c.detach(),
             "psnr_c": utils.compute_psnr(loss_c.detach()), "weight_l2": weight_l2.detach(),
             "loss_bg": (args.bg_weight * loss_bg).detach() if torch.is_tensor(loss_bg) else loss_bg,
             "loss_bg_smooth": loss_bg_smooth.detach() if torch.is_tensor(loss_bg_smooth) else loss_bg_smooth,
             "loss_sp": torch.zeros((), device=rgb.device), "loss_nrm": torch.zeros((), device=rgb.device),
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


def train_step(model, rng, state: TrainState, batch: Dict, args=None, world_size: int = 1, group=None,
               jitter=None, u=None):
    """One optimisation step (train.py:58-183)."""
    args = args if args is not None else batch["args"]
    key_0, key_1 = utils._split_key(rng)
    state.opt.zero_grad(set_to_none=True)
    total, stats = loss_fn(model, state.params, batch, args, key_0, key_1, jitter=jitter, u=u)
    total.backward()
    allreduce_mean_grads(state.params, world_size, group)
    if world_size > 1:
        import torch.distributed as dist
        keys = [k for k, v in stats.items() if torch.is_tensor(v)]
        packed = torch.stack([stats[k].float() for k in keys])
        dist.all_reduce(packed, group=group)                      # pmean(stats), train.py:167
        for k, v in zip(keys, packed / world_size):
            stats[k] = v
    if args.grad_max_val > 0:
        for p in tree_leaves(state.params):
            if p.grad is not None:
                p.grad.clamp_(-args.grad_max_val, args.grad_max_val)
    if args.grad_max_norm > 0:
        gs = [p.grad for p in tree_leaves(state.params) if p.grad is not None]
        norm = torch.sqrt(sum((g ** 2).sum() for g in gs))
        mult = torch.clamp(args.grad_max_norm / (1e-7 + norm), max=1.0)
        for g in gs:
            g.mul_(mult)
    lr = utils.learning_rate_decay(state.step, args.lr_init, args.lr_final, args.max_steps, args.lr_delay_steps,
                                   args.lr_delay_mult)
    for gparam in state.opt.param_groups:
        gparam["lr"] = lr
    state.opt.step()
    state.step += 1
    stats["lr"] = lr
    return state, stats, (int(rng) + 1 if rng is not None else 1)
