"""Development aid: march-kernel experiments on the ship workload (what limits it?)."""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops, synthetic, utils  # noqa: E402
ap = argparse.ArgumentParser(); ap.add_argument("--grid", type=int, default=512); a = ap.parse_args()
G = a.grid
ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
data = synthetic.ellipsoid_occupancy(G, 1.5, (1.0, 0.4, 0.6), ss=4)
n = ops.grid_blur(synthetic.rescale_ior(data, "ship"), ndim, 9, 3.0)
table = ops.grid_table(n.reshape(-1), ndim, nmin, nmax); bricks = ops.grid_bricks(table, ndim)
rays = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), 800, 800)
flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]).cuda(), rays)
S = 768
def run(B, compact, dbg, label):
    os.environ["RNERF_MARCH_DEBUG"] = str(dbg)
    r0 = (640000 - B) // 2
    o = flat.origins[r0:r0 + B].contiguous(); d = flat.viewdirs[r0:r0 + B].contiguous()
    out = ops.BentPath(torch.empty(B, S, 8 if compact else 12, device="cuda"), torch.empty(B, S, device="cuda"))
    f = lambda: ops.march(table, ndim, nmin, nmax, o, d, 2.0, 6.0, S, out=out, bricks=bricks, compact=compact)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    e[0].record()
    for i in range(5):
        f(); e[i + 1].record()
    torch.cuda.synchronize()
    ms = min(e[i].elapsed_time(e[i + 1]) for i in range(5))
    print(f"{label:34s} B={B:7d} {ms:8.3f} ms  {B * (24 + 44 * S) / ms / 1e6:6.0f} GB/s algorithmic  {ms * 1e6 / B:6.2f} ns/ray")
    del out
for B in (128000, 256000, 640000):
    run(B, True, 0, "compact")
run(128000, True, 1, "compact, no record stores")
run(128000, True, 3, "compact, no stores at all")
run(128000, True, 2, "compact, no t-column stores")
run(128000, True, 4, "compact, plain (non-.cs) stores")
run(128000, False, 0, "full")
run(128000, False, 2, "full, no t-column stores")
run(640000, False, 0, "full")
