"""Development aid: tensor-pipe so3 evaluator vs the CUDA-core chain (accuracy + time per 64-point evaluation)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops  # noqa: E402
gen = torch.Generator().manual_seed(0)
p = models.init_small_mlp_params(gen, "cuda", in_dim=60, out_std=0.05)
for d in p.values():
    d["bias"].copy_(((torch.rand(d["bias"].shape, generator=gen) * 2 - 1) * 0.05).cuda())
w = ops.so3_pack(p)
packed = ops.so3_tc_pack(w)
window = [1.0] * 7 + [0.5, 0.0, 0.0]
for N in (1, 63, 64, 65, 1000, 148 * 64 * 40):
    pts = ((torch.rand(N, 3, generator=gen) * 2 - 1) * 1.5).cuda()
    cond = torch.randn(N, 3, generator=gen).cuda()
    a = ops.so3_predict(w, window, pts, cond)
    b = ops.so3_predict_tc(packed, w, window, pts, cond)
    torch.cuda.synchronize()
    err = (a - b).abs().max().item()
    print(f"N={N}: max |tc - cuda-core| = {err:.3e}  (|pred| max {a.abs().max().item():.3f})  finite {bool(torch.isfinite(b).all())}")
def t_ms(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ta = t_ms(lambda: ops.so3_predict(w, window, pts, cond)); tb = t_ms(lambda: ops.so3_predict_tc(packed, w, window, pts, cond))
evals = N / 64
print(f"cuda-core {ta:.3f} ms ({ta * 1e3 * 148 / evals:.2f} us per 64-point evaluation per SM), tensor-pipe {tb:.3f} ms ({tb * 1e3 * 148 / evals:.2f} us)")

for dbg in ("1", "2", "4", "8", "3", "15", "0"):
    os.environ["RNERF_SO3_TC_DEBUG"] = dbg
    tb = t_ms(lambda: ops.so3_predict_tc(packed, w, window, pts, cond))
    print(f"RNERF_SO3_TC_DEBUG={dbg} (1 no MMA, 2 no epilogue, 4 no encoding, 8 no head): {tb * 1e3 * 148 / evals:.2f} us per evaluation")
