"""Where the "all"-stage training march spends its time on a batch of random pixels: active march steps per ray, and the
number of MLP evaluations each CTA of the so3 kernels has to run in series (the union of its rays' active steps), for the
caller's ray order and for the orders autograd._activity_order could use.  python scripts/all_stage_batch_probe.py [--rays N]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import autograd as ag, models, ops, synthetic, utils  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--rays", type=int, default=4096); ap.add_argument("--grid", type=int, default=512)
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
idx = torch.randint(0, 640000, (a.rays,), generator=torch.Generator().manual_seed(0))
o = flat.origins[idx].to(dev).contiguous(); d = flat.viewdirs[idx].to(dev).contiguous()
full = ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=model.bricks, compact=False)
act = full.rec[..., 8:11].norm(dim=-1) > 1e-3                      # [B, S]
per_ray = act.sum(1)
hit = per_ray > 0
print(f"rays {a.rays}: {hit.float().mean().item() * 100:.1f} % meet the boundary; active steps per such ray: mean "
      f"{per_ray[hit].float().mean().item():.1f}, max {per_ray.max().item()}")
S = act.shape[1]
k = torch.arange(S, device=dev)
first = torch.where(act, k, S).amin(1); last = torch.where(act, k, -1).amax(1)
orders = {"caller order": torch.arange(a.rays, device=dev),
          "sorted by (first, last) active step": torch.argsort(first * (S + 1) + last + 1),
          "sorted by active-step count": torch.argsort(per_ray, descending=True),
          "sorted by (first step / 8, count)": torch.argsort((first // 8) * 1024 + (1023 - per_ray.clamp(max=1023)))}
for q in (4, 8, 16, 32, 64):
    orders[f"sorted by (first / {q}, last)"] = torch.argsort((first // q) * (S + 1) + last + 1)
    orders[f"sorted by (last / {q}, first)"] = torch.argsort(((last + 1) // q) * (S + 1) + first)


def morton(x, y):
    z = torch.zeros_like(x)
    for b in range(10):
        z |= ((x >> b) & 1) << (2 * b + 1)
        z |= ((y >> b) & 1) << (2 * b)
    return z


orders["Morton order of (first, last)"] = torch.argsort(morton(first.clamp(max=1023), (last + 1).clamp(max=1023)))
for rpc in (16, 32):
    for name, perm in orders.items():
        u = act[perm][: a.rays // rpc * rpc].reshape(-1, rpc, S).any(1).sum(1)        # evaluations per CTA
        print(f"  {rpc:3d} rays/CTA, {name:38s}: evaluations per CTA max {u.max().item():4d}, mean {u.float().mean().item():6.1f}, "
              f"total {u.sum().item()}")
