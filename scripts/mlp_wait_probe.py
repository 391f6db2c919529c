"""Development aid (timing experiments; results are numerically WRONG when RNERF_PAIR_DEBUG is set): how much of the
CTA-pair enc+MLP kernel's time is the weight-barrier waits on the MMA issuer's critical path, and what the L2 fetch
granularity does to the gather kernels (select / resample)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops, synthetic, utils  # noqa: E402

M = 640000 * 64
gen = torch.Generator().manual_seed(0)
p = models.init_nerf_mlp_params(gen, "cuda")
packed = ops.encmlp_pack(p)
pos = (torch.rand(M, 3, device="cuda") * 2 - 1) * 3
d = torch.randn(M, 3, device="cuda"); d = d / d.norm(dim=-1, keepdim=True)


def t_ms(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for dbg in sys.argv[1:] or ["0", "2", "4", "6", "0"]:
    os.environ["RNERF_PAIR_DEBUG"] = dbg
    ms = t_ms(lambda: ops.encmlp_fwd(packed, pos, d))
    print(f"RNERF_PAIR_DEBUG={dbg}: {ms:.2f} ms  {2 * 593408 * M / ms / 1e9:.0f} TFLOP/s")
os.environ["RNERF_PAIR_DEBUG"] = "0"
del pos, d

# ---- L2 fetch granularity vs the gather kernels
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
G = 512
ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
data = synthetic.ellipsoid_occupancy(G, 1.5, (1.0, 0.4, 0.6), ss=4)
n = ops.grid_blur(synthetic.rescale_ior(data, "ship"), ndim, 9, 3.0)
args = utils.Flags(config="ship_skydome", num_path_samples=12, white_bkgd=False, use_online_sparsity=False)
model, variables = models.construct_nerf(0, None, args, ndim, nmin, nmax, n)
rays = utils.generate_rays(synthetic.camera_pose(0.7, 1.0, 4.03), 800, 800, focal=0.5 * 800 / 0.36)
B = 640000
o = rays.origins.reshape(-1, 3).contiguous(); v = rays.viewdirs.reshape(-1, 3).contiguous()
path = ops.march(model.table, ndim, nmin, nmax, o, v, 2.0, 6.0, 768, bricks=model.bricks, compact=True)
jit = model.draw_jitter(1)
pos_c, dir_c, t_c, _ = ops.select(path, jit)
w = torch.rand(B, 64, device="cuda") ** 4
u = model.draw_u(2, B, False)
lim = ctypes.c_size_t(0)
for gran in (None, 32, 64, 128):
    if gran is not None and rt is not None:
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran))
        rt.cudaDeviceGetLimit(ctypes.byref(lim), 5)
        tag = f"set {gran} rc={rc} now={lim.value}"
    else:
        if rt is not None:
            rt.cudaDeviceGetLimit(ctypes.byref(lim), 5)
        tag = f"default ({lim.value})"
    a = t_ms(lambda: ops.select(path, jit))
    b = t_ms(lambda: ops.resample(path, t_c, w, u, 128))
    c = t_ms(lambda: ops.march(model.table, ndim, nmin, nmax, o, v, 2.0, 6.0, 768, out=path, bricks=model.bricks, compact=True))
    print(f"L2 fetch granularity {tag}: select {a:.3f} ms  resample {b:.3f} ms  march {c:.3f} ms")
