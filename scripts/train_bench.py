"""Secondary benchmark (BASELINE.json configs[2]): ship_skydome training step, 4096-ray batch per GPU step (forward +
backward + bucketed gradient all-reduce + Adam).  Prints one JSON line; run under torchrun for N > 1.
Not the headline metric (bench.py is); timed like it: W warm-up steps, K timed steps between barriers, CUDA events."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import _lib, models, ops, synthetic, train, utils  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=4096, help="global batch (rays per step over all GPUs)")
ap.add_argument("--grid", type=int, default=512)
ap.add_argument("--eager", action="store_true", help="no CUDA-graph capture of the step")
ap.add_argument("--stage", default="radiance", help='"radiance" (configs[2]) or "all" (so3_mlp trained through the scan adjoint)')
ap.add_argument("--learn-grid", action="store_true", help="extension: the IoR grid is a trainable parameter too (gradient by the "
                "reverse sweep, all-reduced and updated by a fused Adam)")
ap.add_argument("--emulate-world", type=int, default=0, help="development aid: one process with the per-rank shapes of a "
                "W-GPU step (batch / W rays, 128 / W env rows) and no communicator -- the per-rank latency floor of the step")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
G = a.grid
ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
data = synthetic.ellipsoid_occupancy(G, 1.5, (1.0, 0.4, 0.6), ss=4, device=dev)
n = ops.grid_blur(synthetic.rescale_ior(data, "ship_skydome"), ndim, 9, 3.0)
del data
args = utils.Flags(config="ship_skydome-bkgd_no-partial-reflect_cycles", num_path_samples=12, white_bkgd=False,
                   use_online_sparsity=False, bg_weight=0.025, bg_smooth_weight=1.0, bg_patch_size=128, randomized=True,
                   max_steps=200000, lr_delay_steps=2500, lr_delay_mult=0.01, stage=a.stage)
model, variables = models.construct_nerf(0, None, args, ndim, nmin, nmax, n)
if a.learn_grid:
    model.enable_grid_learning()
state = train.TrainState.create(variables, args, model=model)
shape_world = a.emulate_world if a.emulate_world > 0 else world
B = a.batch // shape_world
rays_hw = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), 800, 800)
flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]), rays_hw)
gen = torch.Generator().manual_seed(rank)
idx = torch.randint(0, 640000, (B,), generator=gen)
rays = utils.namedtuple_map(lambda r: r[idx].to(dev).contiguous(), flat)
env = synthetic.blender_rays(synthetic.camera_pose(1.3, 0.8, 4.03), 128, 128, camera_angle_x=0.2)
_e0, _e1 = utils.shard_range(128, rank, shape_world)       # the reference shards the whole batch dict, env patch included (utils.shard)
env = utils.namedtuple_map(lambda r: r[_e0:_e1].to(dev).contiguous(), env)
batch = {"rays": rays, "pixels": torch.rand(B, 3, generator=gen).to(dev), "env_rays": env, "annealed_alpha": 0.5}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


rng = 0
state.step = 3000
for _ in range(a.warmup):
    state, stats, rng = train.train_step(model, rng, state, batch, args, world_size=world, use_graph=False if a.eager else None)
barrier()
l0 = _lib.launch_count() + state.replayed_kernel_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
torch.cuda.cudart().cudaProfilerStart()
tc0 = time.perf_counter()
e0.record()
for _ in range(a.steps):
    state, stats, rng = train.train_step(model, rng, state, batch, args, world_size=world, use_graph=False if a.eager else None)
e1.record()
cpu_issue_ms = (time.perf_counter() - tc0) * 1e3 / a.steps
barrier()
torch.cuda.cudart().cudaProfilerStop()
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
if rank == 0:
    print(json.dumps({"metric": "training rays/sec (fwd+bwd+allreduce+Adam)", "value": a.batch * a.steps / (ms * 1e-3),
                      "unit": "rays/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
                      "scaling": "strong", "global_batch": a.batch, "stage": a.stage, "learn_grid": a.learn_grid, "cuda_graph": not a.eager, "gpu_launches": int(_lib.launch_count() + state.replayed_kernel_launches() - l0),
                      "loss": float(stats["loss"]), "cpu_issue_ms_per_step": cpu_issue_ms, "config": "ship_skydome training step, S=768, G=%d, 64+192 samples, "
                      "bg_weight 0.025, bg_smooth 1.0 on a 128x128 env patch" % G}))
if world > 1:
    train.shutdown_distributed(state)
