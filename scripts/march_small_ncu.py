"""Radiance-stage march of a 4096-ray training batch (random pixels), for ncu: the launch is latency-bound (32 CTAs)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops, synthetic, utils  # noqa: E402
dev = torch.device("cuda", 0)
G = 512
ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
data = synthetic.ellipsoid_occupancy(G, 1.5, (1.0, 0.4, 0.6), ss=4, device=dev)
n = ops.grid_blur(synthetic.rescale_ior(data, "ship_skydome"), ndim, 9, 3.0)
table = ops.grid_table(n.reshape(-1), ndim, nmin, nmax); bricks = ops.grid_bricks(table, ndim)
rays = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), 800, 800)
flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]), rays)
idx = torch.randint(0, 640000, (4096,), generator=torch.Generator().manual_seed(0))
o = flat.origins[idx].to(dev).contiguous(); d = flat.viewdirs[idx].to(dev).contiguous()
path = ops.march(table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=bricks, compact=True)
for _ in range(3):
    ops.march(table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=bricks, compact=True, out=path)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.march(table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=bricks, compact=True, out=path)
e1.record(); torch.cuda.synchronize()
print("march 4096 rays: %.3f ms" % (e0.elapsed_time(e1) / 10))
