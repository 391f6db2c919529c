"""N-GPU training check (run under torchrun on a GPU box): the all-reduce-mean of per-shard gradients equals the
single-GPU gradient of the concatenated batch (bg_weight = 0 so that the per-shard normaliser of loss_bg, which the
reference also has, does not enter), and every rank ends the step with identical parameters."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from samplenerfro_b200 import models, train, utils, synthetic, ops  # noqa: E402

STAGE = sys.argv[1] if len(sys.argv) > 1 else "radiance"        # "radiance" or "all" (so3_mlp trained through the scan adjoint)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G = 32
ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
data = synthetic.ellipsoid_occupancy(G, 1.5, (0.8, 0.6, 0.7), ss=2)
n = ops.grid_blur(synthetic.rescale_ior(data, "example"), ndim, 3, 1.0)
args = utils.Flags(config="example", num_path_samples=12, white_bkgd=False, use_online_sparsity=False, bg_weight=0.0,
                   bg_smooth_weight=0.0, randomized=True, max_steps=1000, lr_delay_steps=0, stage=STAGE)
B = 64 * world
gen = torch.Generator().manual_seed(0)
rays_hw = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), 16, 16)
flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1])[:B].cuda(), rays_hw)
pixels = torch.rand(B, 3, generator=gen).cuda()
u = torch.clamp(torch.arange(128) / 128 + torch.rand(B, 128, generator=gen) * (1 / 128 - 1e-7), max=1 - 1.2e-7).cuda()


def grads_of(rays, px, uu, ws):
    model, variables = models.construct_nerf(3, None, args, ndim, nmin, nmax, n)
    if STAGE == "all":        # a rotation large enough to matter (the init is N(0, 1e-5))
        g4 = torch.Generator().manual_seed(11)
        d4 = variables["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]["Dense_4"]
        d4["kernel"].copy_((torch.randn(128, 3, generator=g4) * 0.05).cuda()); d4["bias"].copy_((torch.randn(3, generator=g4) * 0.2).cuda())
    state = train.TrainState.create(variables, args)
    batch = {"rays": rays, "pixels": px, "annealed_alpha": 0.5}
    total, _ = train.loss_fn(model, variables, batch, args, 1, 2, jitter=model.draw_jitter(5), u=uu)
    total.backward()
    state.arena.allreduce_mean(ws)
    return variables, state, model, batch


lo, hi = utils.shard_range(B, rank, world)
shard = utils.namedtuple_map(lambda r: r[lo:hi].contiguous(), flat)
v_dist, state, model, batch = grads_of(shard, pixels[lo:hi].contiguous(), u[lo:hi].contiguous(), world)
v_full, _, _, _ = grads_of(flat, pixels, u, 1)
worst = 0.0
for name in state.arena.buckets:
    for a, b in zip(train.tree_leaves(v_dist["params"][name]), train.tree_leaves(v_full["params"][name])):
        rel = ((a.grad - b.grad).norm() / (b.grad.norm() + 1e-20)).item()
        worst = max(worst, rel)
# one optimiser step, then every rank must hold the same parameters
state, stats, _ = train.train_step(model, 0, state, batch, args, world_size=world, jitter=model.draw_jitter(5), u=u[lo:hi].contiguous())
flatp = torch.cat([p.detach().reshape(-1) for p in train.tree_leaves(state.params["params"]["fine_mlp"])])
ref = flatp.clone(); dist.broadcast(ref, 0)
same = torch.equal(ref, flatp)
if rank == 0:
    print(f"stage={STAGE} buckets={state.arena.buckets}")
    print(f"world={world} worst relative gradient difference vs single-GPU full batch: {worst:.3e}; params identical after step: {same}")
    assert worst < 2e-2, worst      # bf16 forward: shard order changes tile composition, not the math
assert same
# graph mode: two eager steps, then the captured step (NCCL all-reduces inside the CUDA graph) replayed three times
for i in range(5):
    state, stats, _ = train.train_step(model, i + 1, state, batch, args, world_size=world)
torch.cuda.synchronize()
captured = any(isinstance(v, train._GraphedStep) for v in state.graphs.values())
ref = state.arena.theta.clone(); dist.broadcast(ref, 0)
same_g = torch.equal(ref, state.arena.theta)
loss = float(stats["loss"])
if rank == 0:
    print(f"graph-captured step with in-graph NCCL: captured={captured}, params identical on all ranks: {same_g}, loss {loss:.5f}")
assert captured and same_g and np.isfinite(loss)
train.shutdown_distributed(state)
