"""Secondary configurations of BASELINE.json, each printed as JSON lines (not the headline; bench.py is):
  --config ball   configs[3]: ball.gin shape -- 1008x756 OpenCV camera, S=1536 (P=24), near/far 0.2/12, G=256 extent 2,
                  blur 5/3, bd_cut_dist=6 with the hard-coded ball box; rays sharded over the ranks (row bands)
  --config sweep  configs[4]: march microbenchmark, B in 2^16..2^24 x S in {64..512} x G in {128,256,512} vs HBM roofline
Run under torchrun for N > 1 (ball).  CUDA events, warm-up, max over ranks like bench.py."""
import argparse, json, math, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops, synthetic, utils  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="ball", choices=["ball", "sweep"])
ap.add_argument("--steps", type=int, default=3); ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--quick", action="store_true", help="sweep: a reduced set of points")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
HBM = float(peaks.get("hbm_gbs", 6500.0))


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
    return ms / steps


if a.config == "ball":
    G = 256
    ndim, nmin, nmax = [G] * 3, [-2.0] * 3, [2.0] * 3
    data = synthetic.ellipsoid_occupancy(G, 2.0, (1.0, 1.0, 1.0), center=(0.0, 1.036, 0.0), ss=4, device=dev)
    n = ops.grid_blur(synthetic.rescale_ior(data, "ball"), ndim, 5, 3.0)
    flags = utils.Flags(config="ball", num_path_samples=24, white_bkgd=False, use_online_sparsity=False, near=0.2, far=12.0)
    flags.gin_bindings = {"NerfModel": {"bd_cut_dist": 6.0}}
    model, variables = models.construct_nerf(0, None, flags, ndim, nmin, nmax, n)
    H, W = 756, 1008
    # OpenCV camera 5 units in front of the ball, looking at it (+z forward, +y down in camera space)
    c2w = np.eye(4); c2w[:3, 3] = [0.0, 1.036, -5.0]
    K = np.array([[1100.0, 0, W / 2], [0, 1100.0, H / 2], [0, 0, 1]])
    out = {}
    render_fn = lambda k0, k1, rays: model.apply(variables, k0, k1, rays, False)
    per = -(-H // world)

    def frame():
        # every rank generates and renders its row band; the bands are all-gathered into the full frame on every rank
        out["frame"] = utils.render_view_sharded(render_fn, c2w, H, W, 0, cam_mat=K, chunk=per * W, rank=rank, world_size=world,
                                                 device=dev)

    with torch.no_grad():
        ms = timed(frame, a.steps, a.warmup)
        diff = None
        if world > 1:       # the assembled frame against one GPU rendering all rows (tile composition of the MLP kernel differs)
            single = utils.render_view(render_fn, c2w, H, W, 0, cam_mat=K, chunk=per * W, device=dev)
            diff = max((x - y).abs().max().item() for x, y in zip(out["frame"], single))
    if rank == 0:
        print(json.dumps({"metric": "rays/sec (march+MLP+composite)", "value": H * W / (ms * 1e-3), "unit": "rays/s", "n_gpus": world,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "scaling": "strong",
                          "sharded_vs_single_gpu_max_abs_diff": diff,
                          "config": {"workload": "ball.gin shape: 1008x756 OpenCV view, S=1536 eikonal steps, IoR grid 256^3 "
                                     "(extent 2), 64 + 192 MLP samples/ray, bd_cut_dist=6 (2 extra composites), rays generated on "
                                     "the device, row bands per rank, bands all-gathered into the frame on every rank (inside the "
                                     "timed region)", "rays_per_step": H * W}}), flush=True)
else:
    res = []
    Gs = [512] if a.quick else [128, 256, 512]
    for G in Gs:
        ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
        data = synthetic.ellipsoid_occupancy(G, 1.5, (0.5, 0.5, 0.5), ss=2, device=dev)
        n = ops.grid_blur(synthetic.rescale_ior(data, "sweep"), ndim, 3, 1.0)
        table = ops.grid_table(n.reshape(-1), ndim, nmin, nmax)
        bricks = ops.grid_bricks(table, ndim)
        gen = torch.Generator(device=dev).manual_seed(0)
        for logB in ([20] if a.quick else [16, 18, 20, 22, 24]):
            B = 1 << logB
            # SURVEY 8(d) config E rays: origins on a radius-4 sphere, directions towards uniform points of the unit ball
            oo = torch.nn.functional.normalize(torch.randn(B, 3, generator=gen, device=dev), dim=-1) * 4.0
            tgt = torch.nn.functional.normalize(torch.randn(B, 3, generator=gen, device=dev), dim=-1) * \
                torch.rand(B, 1, generator=gen, device=dev) ** (1 / 3)
            dd = torch.nn.functional.normalize(tgt - oo, dim=-1)
            # sort rays so that a warp holds neighbouring rays, like adjacent pixels of an image do
            key = (torch.atan2(oo[:, 1], oo[:, 0]) * 64).floor() * 1e4 + torch.acos((oo[:, 2] / 4).clamp(-1, 1)) * 1e3
            order = torch.argsort(key)
            oo, dd = oo[order].contiguous(), dd[order].contiguous()
            for S in ([256] if a.quick else [64, 128, 256, 512]):
                chunk = min(B, 1 << int(math.log2(max(1 << 16, (1 << 33) // (S * 36)))))      # power of two, <= 8 GiB of path per launch
                path = ops.BentPath(torch.empty(chunk, S, 8, device=dev), torch.empty(chunk, S, device=dev))

                def run():
                    for i in range(0, B, chunk):
                        ops.march(table, ndim, nmin, nmax, oo[i:i + chunk], dd[i:i + chunk], 2.0, 6.0, S, out=path,
                                  bricks=bricks, compact=True)
                ms = timed(run, 2, 1)
                alg = B * (24 + 44 * S)
                res.append({"G": G, "log2_rays": logB, "S": S, "ms": ms, "algorithmic_GBps": alg / ms / 1e6,
                            "frac_of_hbm_peak": alg / ms / 1e6 / HBM, "rays_per_s": B / (ms * 1e-3)})
                print(json.dumps(res[-1]), flush=True)
                del path
        del table, bricks
    print(json.dumps({"summary": "march sweep", "hbm_peak_GBps": HBM, "points": len(res),
                      "median_frac": sorted(r["frac_of_hbm_peak"] for r in res)[len(res) // 2],
                      "max_frac": max(r["frac_of_hbm_peak"] for r in res)}))
if world > 1:
    dist.destroy_process_group()
