"""Timing of the "all"-stage march (so3_mlp inside the eikonal steps, a4) against the radiance-stage march on the ship
workload (S=768, G=512): python scripts/all_stage_probe.py [--rays N]."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops, synthetic, utils  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--rays", type=int, default=65536); ap.add_argument("--grid", type=int, default=512)
a = ap.parse_args()
dev = torch.device("cuda", 0)
G = a.grid
ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
data = synthetic.ellipsoid_occupancy(G, 1.5, (1.0, 0.4, 0.6), ss=4, device=dev)
n = ops.grid_blur(synthetic.rescale_ior(data, "ship_skydome"), ndim, 9, 3.0)
args = utils.Flags(config="ship_skydome", num_path_samples=12, white_bkgd=False, use_online_sparsity=False, stage="all")
model, variables = models.construct_nerf(0, None, args, ndim, nmin, nmax, n)
rays = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), 800, 800)
flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]), rays)
idx = torch.linspace(0, 640000 - 1, a.rays).long()
o = flat.origins[idx].to(dev).contiguous(); d = flat.viewdirs[idx].to(dev).contiguous()
so3 = (model._so3_packed(variables), model.so3_window(1.0))


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


path = ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=model.bricks, compact=True)
t_rad = timeit(lambda: ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=model.bricks, compact=True, out=path))
t_all = timeit(lambda: ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=model.bricks, compact=True, out=path, so3=so3))
full = ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=model.bricks, compact=False)
g = full.rec[..., 8:11].norm(dim=-1)
act = (g > 1e-3).float()
warp_act = act.reshape(-1, 32, 768).amax(dim=1)
print(f"rays {a.rays}: radiance march {t_rad:.3f} ms, all-stage march {t_all:.3f} ms; steps with |grad n| > 1e-3: "
      f"{100 * act.mean().item():.2f} % of ray-steps, {100 * warp_act.mean().item():.2f} % of warp-steps "
      f"({act.sum().item() * 2 * 64896 / t_all / 1e9:.2f} TFLOP/s of useful so3 flops)")
