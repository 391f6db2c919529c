"""Timing of the "all"-stage march (so3_mlp inside the eikonal steps, a4) against the radiance-stage march on the ship
workload (S=768, G=512): python scripts/all_stage_probe.py [--rays N]."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops, synthetic, utils  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--rays", type=int, default=640000); ap.add_argument("--grid", type=int, default=512)
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
r0 = (640000 - a.rays) // 2          # a contiguous band of image rows: a warp / CTA holds adjacent pixels
o = flat.origins[r0:r0 + a.rays].to(dev).contiguous(); d = flat.viewdirs[r0:r0 + a.rays].to(dev).contiguous()
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
tc = model._so3_tc_packed(variables)
t_tc = timeit(lambda: ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=model.bricks, compact=True, out=path, so3=so3, so3_tc=tc))
print(f"all-stage march on the tensor pipe (march_tc_kernel): {t_tc:.3f} ms  (CUDA-core chain {t_all:.3f} ms, radiance {t_rad:.3f} ms)")
for dbg in os.environ.get("PROBE_DBG", "").split():
    os.environ["RNERF_SO3_TC_DEBUG"] = dbg
    tt = timeit(lambda: ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, 768, bricks=model.bricks, compact=True, out=path, so3=so3, so3_tc=tc))
    print(f"  RNERF_SO3_TC_DEBUG={dbg} (1 no MMA, 2 no epilogue, 4 no encoding, 8 no head, 16 no evaluation): {tt:.3f} ms")
os.environ["RNERF_SO3_TC_DEBUG"] = "0"
del path
n_ray = n_warp = n_cta = 0.0
for i in range(0, a.rays // 128 * 128, 65536):
    j = min(i + 65536, a.rays // 128 * 128)
    full = ops.march(model.table, ndim, nmin, nmax, o[i:j], d[i:j], 2.0, 6.0, 768, bricks=model.bricks, compact=False)
    act = (full.rec[..., 8:11].norm(dim=-1) > 1e-3).float()
    n_ray += act.sum().item()
    n_warp += act.reshape(-1, 32, 768).amax(dim=1).sum().item()
    n_cta += act.reshape(-1, 128, 768).amax(dim=1).sum().item()
    n_cta256 = globals().get("n_cta256", 0.0) + act.reshape(-1, 256, 768).amax(dim=1).sum().item()
    n_over64 = globals().get("n_over64", 0.0) + (act.reshape(-1, 256, 768).sum(dim=1) > 64).float().sum().item()
    del full, act
tot = a.rays * 768
print(f"256-ray CTAs: {n_cta256:.0f} CTA-steps need an evaluation, {n_over64:.0f} of them more than 64 columns")
print(f"rays {a.rays}: radiance march {t_rad:.3f} ms, all-stage march {t_all:.3f} ms; |grad n| > 1e-3 at {100 * n_ray / tot:.2f} % of "
      f"ray-steps, {100 * n_warp * 32 / tot:.2f} % of warp-steps, {100 * n_cta * 128 / tot:.2f} % of CTA-steps = {n_cta:.0f} CTA "
      f"evaluations -> {(t_all - t_rad) * 1e3 * 148 / max(n_cta, 1):.1f} us per evaluation per SM; "
      f"{n_ray * 2 * 64896 / t_all / 1e9:.2f} TFLOP/s of useful so3 flops")
