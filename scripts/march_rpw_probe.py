"""Development aid: radiance-stage march on small launches (training batches of random pixels) against the number of rays a
warp carries (RNERF_MARCH_RPW); checks the records are bit-identical to the 32-rays-per-warp launch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import ops, synthetic, utils  # noqa: E402
G = 512
ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
data = synthetic.ellipsoid_occupancy(G, 1.5, (1.0, 0.4, 0.6), ss=4)
n = ops.grid_blur(synthetic.rescale_ior(data, "ship"), ndim, 9, 3.0)
table = ops.grid_table(n.reshape(-1), ndim, nmin, nmax); bricks = ops.grid_bricks(table, ndim)
rays = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), 800, 800)
flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]).cuda(), rays)
S = 768
gen = torch.Generator().manual_seed(0)
for B in (500, 512, 4096, 16384, 65536, 262144):
    idx = torch.randint(0, 640000, (B,), generator=gen).cuda()
    o = flat.origins[idx].contiguous(); d = flat.viewdirs[idx].contiguous()
    ref = None
    for rpw in ("32", "16", "8", "4", "2", "1", None):
        if rpw is None:
            os.environ.pop("RNERF_MARCH_RPW", None)
        else:
            os.environ["RNERF_MARCH_RPW"] = rpw
        out = ops.BentPath(torch.zeros(B, S, 8, device="cuda"), torch.zeros(B, S, device="cuda"))
        f = lambda: ops.march(table, ndim, nmin, nmax, o, d, 2.0, 6.0, S, out=out, bricks=bricks, compact=True)
        for _ in range(2): f()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        e[0].record()
        for i in range(5):
            f(); e[i + 1].record()
        torch.cuda.synchronize()
        ms = min(e[i].elapsed_time(e[i + 1]) for i in range(5))
        if ref is None:
            ref = (out.rec.clone(), out.t.clone())
        same = torch.equal(out.rec.view(torch.int32), ref[0].view(torch.int32)) and torch.equal(out.t.view(torch.int32), ref[1].view(torch.int32))
        print(f"B={B:7d} rpw={str(rpw):>4s} {ms:8.3f} ms  bit-identical={same}", flush=True)
