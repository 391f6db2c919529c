"""Development aid: time the CTA-pair enc+MLP kernel under the RNERF_PAIR_DEBUG timing experiments."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops  # noqa: E402
M = 128000 * 192
gen = torch.Generator().manual_seed(0)
p = models.init_nerf_mlp_params(gen, "cuda")
packed = ops.encmlp_pack(p)
pos = (torch.rand(M, 3, device="cuda") * 2 - 1) * 3
d = torch.randn(M, 3, device="cuda"); d = d / d.norm(dim=-1, keepdim=True)
for dbg in [int(x) for x in (sys.argv[1:] or ["0", "1", "3", "7"])]:
    os.environ["RNERF_PAIR_DEBUG"] = str(dbg)
    for _ in range(2): ops.encmlp_fwd(packed, pos, d)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    for i in range(3):
        ops.encmlp_fwd(packed, pos, d); e[i + 1].record()
    torch.cuda.synchronize()
    ms = min(e[i].elapsed_time(e[i + 1]) for i in range(3))
    print(f"dbg={dbg}: {ms:.3f} ms  {2 * 593408 * M / ms / 1e9:.1f} TFLOP/s")
