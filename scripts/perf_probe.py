"""Per-kernel timing probe on the ship-render workload (G=512, S=768, 64+192 samples).  Not a bench: a
development aid that prints CUDA-event times per stage for one chunk of rays."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops, synthetic, utils  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    return min(ts), sum(ts) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--rays", type=int, default=128000)
    ap.add_argument("--side", type=int, default=800)
    a = ap.parse_args()
    G = a.grid
    ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
    t0 = time.time()
    data = synthetic.ellipsoid_occupancy(G, 1.5, (1.0, 0.4, 0.6), ss=4)
    n = ops.grid_blur(synthetic.rescale_ior(data, "ship"), ndim, 9, 3.0)
    torch.cuda.synchronize()
    print(f"grid setup {time.time() - t0:.2f}s  n in [{n.min().item():.3f},{n.max().item():.3f}]")
    args = utils.Flags(config="ship_skydome", num_path_samples=12, white_bkgd=False, use_online_sparsity=False)
    model, variables = models.construct_nerf(0, None, args, ndim, nmin, nmax, n)
    rays = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), a.side, a.side)
    flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]), rays)
    B = a.rays
    r0 = (a.side * a.side - B) // 2
    o = flat.origins[r0:r0 + B].cuda().contiguous(); d = flat.viewdirs[r0:r0 + B].cuda().contiguous()
    S, Nc, Nf = 768, 64, 128
    path = ops.BentPath(torch.empty(B, S, 8, device="cuda"), torch.empty(B, S, device="cuda"))
    path_full = ops.BentPath(torch.empty(B, S, 12, device="cuda"), torch.empty(B, S, device="cuda"))
    jit = model.draw_jitter(1)
    u = model.draw_u(2, B, False)
    pk_c = model._packed(variables, "coarse_mlp"); pk_f = model._packed(variables, "fine_mlp"); wb = model._packed(variables, "bkgd_mlp")
    res = {}
    res["march_full"] = timeit(lambda: ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, S, out=path_full, bricks=model.bricks))
    res["march_nobricks"] = timeit(lambda: ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, S, out=path, compact=True))
    res["march"] = timeit(lambda: ops.march(model.table, ndim, nmin, nmax, o, d, 2.0, 6.0, S, out=path, bricks=model.bricks, compact=True))
    pos_c, dir_c, t_c, _ = ops.select(path, jit)
    res["select"] = timeit(lambda: ops.select(path, jit))
    res["bkgd_mlp"] = timeit(lambda: ops.bkgd_mlp_fwd(wb, dir_c, B, Nc * 3, (Nc - 1) * 3))
    raw_b = ops.bkgd_mlp_fwd(wb, dir_c, B, Nc * 3, (Nc - 1) * 3)
    res["encmlp_coarse"] = timeit(lambda: ops.encmlp_fwd(pk_c, pos_c, dir_c))
    raw_c = ops.encmlp_fwd(pk_c, pos_c, dir_c).view(B, Nc, 4)
    res["composite_c"] = timeit(lambda: ops.composite_fwd(raw_c, t_c, dir_c, raw_b))
    oc = ops.composite_fwd(raw_c, t_c, dir_c, raw_b)
    res["resample"] = timeit(lambda: ops.resample(path, t_c, oc["weights"], u, Nf))
    t_f, pos_f, dir_f, _ = ops.resample(path, t_c, oc["weights"], u, Nf)
    res["encmlp_fine"] = timeit(lambda: ops.encmlp_fwd(pk_f, pos_f, dir_f))
    raw_f = ops.encmlp_fwd(pk_f, pos_f, dir_f).view(B, Nc + Nf, 4)
    res["composite_f"] = timeit(lambda: ops.composite_fwd(raw_f, t_f, dir_f, raw_b, want_weights=False))
    rays_cu = utils.Rays(o, d, d, torch.ones(B, 1, device="cuda"))
    res["model.apply"] = timeit(lambda: model.apply(variables, 1, 2, rays_cu, False), n=3, warm=1)
    tot = 0
    for k, (mn, av) in res.items():
        extra = ""
        if k.startswith("march"):
            wr = B * S * ((48 if k == "march_full" else 32) + 4)
            extra = f"  {B * (24 + 44 * S) / mn / 1e6:.0f} GB/s algorithmic (reference arrays), {wr / mn / 1e6:.0f} GB/s written"
        if k.startswith("encmlp"):
            M = B * (Nc if k.endswith("coarse") else Nc + Nf)
            extra = f"  {2 * 593408 * M / mn / 1e9:.1f} TFLOP/s"
        if k.startswith("composite"):
            Ns = Nc if k.endswith("_c") else Nc + Nf
            extra = f"  {B * (32 * Ns + 36) / mn / 1e6:.0f} GB/s algorithmic"
        if k == "resample":
            extra = f"  {B * (500 + 192 * 80) / mn / 1e6:.0f} GB/s algorithmic"
        print(f"{k:16s} min {mn:9.3f} ms  avg {av:9.3f} ms{extra}")
        if k != "model.apply":
            tot += mn
    print(f"sum of stages {tot:.3f} ms -> {B / tot / 1e3:.3f} M rays/s ; model.apply -> {B / res['model.apply'][0] / 1e3:.3f} M rays/s")


if __name__ == "__main__":
    main()
