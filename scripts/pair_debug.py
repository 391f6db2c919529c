"""Development aid: compare the CTA-pair kernel with the single-CTA kernel row by row."""
import os, sys, subprocess
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 74 * 512 + 100
gen = torch.Generator().manual_seed(0)
p = models.init_nerf_mlp_params(gen, "cuda")
for d in p.values():
    d["bias"].copy_((torch.rand(d["bias"].shape, generator=gen) * 0.2 - 0.1).cuda())
packed = ops.encmlp_pack(p)
pos = (torch.rand(M, 3, generator=gen) * 2 - 1).cuda() * 3
dr = torch.randn(M, 3, generator=gen).cuda(); dr = dr / dr.norm(dim=-1, keepdim=True)
raw_dbg, _ = ops.encmlp_fwd(packed, pos, dr, debug_layers=True)   # single-CTA kernel
raw = ops.encmlp_fwd(packed, pos, dr)                              # pair kernel (M >= 37888)
torch.cuda.synchronize()
err = (raw - raw_dbg).abs()
print("max err", err.max().item(), "scale", raw_dbg.abs().max().item())
bad = (err.max(dim=1).values > 1e-4).nonzero().flatten()
print("bad rows", bad.numel(), "of", M)
if bad.numel():
    print("first bad", bad[:20].tolist())
    r = bad.cpu()
    print("row%512 hist (by 128):", torch.bincount((r % 512) // 128, minlength=4).tolist())
    print("per-channel max err", err.max(dim=0).values.tolist())
    print("group hist first 10:", torch.bincount(r // 512)[:10].tolist())
    print(raw[bad[:4]], raw_dbg[bad[:4]])
