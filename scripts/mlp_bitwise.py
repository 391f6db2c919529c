"""Development aid: checksum of the enc+MLP outputs on fixed inputs (run once per library build with RNERF_LIB=... and
compare the printed digests: a schedule-only change must leave them identical)."""
import hashlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops  # noqa: E402
M = 1 << 20
gen = torch.Generator().manual_seed(0)
p = models.init_nerf_mlp_params(gen, "cuda")
for d in p.values():
    d["bias"].copy_((torch.rand(d["bias"].shape, generator=gen) * 0.2 - 0.1).cuda())
packed = ops.encmlp_pack(p)
pos = ((torch.rand(M, 3, generator=gen) * 2 - 1) * 3).cuda()
dr = torch.randn(M, 3, generator=gen).cuda(); dr = dr / dr.norm(dim=-1, keepdim=True)
raw = ops.encmlp_fwd(packed, pos, dr)
raw2, saved = ops.encmlp_fwd_train(packed, pos, dr)
torch.cuda.synchronize()
print("fwd", hashlib.sha256(raw.cpu().numpy().tobytes()).hexdigest()[:16], "train", hashlib.sha256(raw2.cpu().numpy().tobytes()).hexdigest()[:16],
      "layers", hashlib.sha256(saved[0].view(torch.int16).cpu().numpy().tobytes()).hexdigest()[:16], "finite", bool(torch.isfinite(raw).all()))
