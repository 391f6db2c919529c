"""Development aid: per-layer cycle breakdown of the enc+MLP kernel (CTA 0, first tile group)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536 * 64
gen = torch.Generator().manual_seed(0)
p = models.init_nerf_mlp_params(gen, "cuda")
packed = ops.encmlp_pack(p)
pos = (torch.rand(M, 3, device="cuda") * 2 - 1) * 3
d = torch.randn(M, 3, device="cuda"); d = d / d.norm(dim=-1, keepdim=True)
for _ in range(2):
    raw, prof = ops.encmlp_fwd_profile(packed, pos, d)
torch.cuda.synchronize()
pr = prof.cpu()
t0 = pr[0, 0, 0].item()
print("layer | MMA: wait_A  issue   | EPI: wait_acc  work  | (cycles)   abs start")
for l in range(10):
    m = pr[0, l]; e = pr[1, l]
    print(f"{l:5d} | {m[1]-m[0]:8d} {m[2]-m[1]:8d} | {e[1]-e[0]:8d} {e[2]-e[1]:8d} | mma_start {m[0]-t0:8d} epi_done {e[2]-t0:8d}")
print("group total cycles:", (pr[1, 9, 2] - t0).item())

