"""One forward "all"-stage march and one reverse sweep on a band of image rows (or random pixels with --random), for ncu:
  ncu --set full --import-source on --clock-control none -k regex:'march_kernel|march_all_bwd' -o out python scripts/all_stage_ncu.py"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops, synthetic, utils  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--rays", type=int, default=131072); ap.add_argument("--grid", type=int, default=512)
ap.add_argument("--random", action="store_true", help="random pixels (a training batch) instead of a band of rows")
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
if a.random:
    idx = torch.randint(0, 640000, (a.rays,), generator=torch.Generator().manual_seed(0))
else:
    r0 = (640000 - a.rays) // 2
    idx = torch.arange(r0, r0 + a.rays)
o = flat.origins[idx].to(dev).contiguous(); d = flat.viewdirs[idx].to(dev).contiguous()
so3 = (model._so3_packed(variables), model.so3_window(1.0))
path = ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=model.bricks, compact=True, so3=so3,
                 so3_tc=None if a.random else model._so3_tc_packed(variables))      # band of rows: the tensor-pipe kernel
jitter = torch.arange(0, 768, 12, dtype=torch.int32, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)
gp = torch.randn(a.rays, 64, 3, device=dev, generator=gen); gd = torch.randn(a.rays, 64, 3, device=dev, generator=gen)
g, _, _ = ops.march_all_bwd(model.table, ndim, nmin, nmax, path, 2.0, 6.0, jitter, gp, gd, so3, bricks=model.bricks)
torch.cuda.synchronize()
print("ok", float(g.abs().sum()))
