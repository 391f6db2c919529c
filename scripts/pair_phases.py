"""Development aid: per-layer cycle stamps of the CTA-pair enc+MLP kernel (leader CTA 0, second tile group)."""
import os, sys
os.environ["RNERF_PROFILE_PAIR"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplenerfro_b200 import models, ops  # noqa: E402

M = 65536 * 64
gen = torch.Generator().manual_seed(0)
p = models.init_nerf_mlp_params(gen, "cuda")
packed = ops.encmlp_pack(p)
pos = (torch.rand(M, 3, device="cuda") * 2 - 1) * 3
d = torch.randn(M, 3, device="cuda"); d = d / d.norm(dim=-1, keepdim=True)
for _ in range(2):
    raw, prof = ops.encmlp_fwd_profile(packed, pos, d)
torch.cuda.synchronize()
pr = prof.cpu().reshape(-1)
mma = pr[:80].reshape(10, 2, 4); epi = pr[80:160].reshape(10, 2, 4)
t0 = mma[0, 0, 0].item()
print("layer pair | MMA: wait_A  issue | start_abs | EPI: wait_acc  work  done_abs")
for l in range(10):
    for j in range(2):
        m = mma[l, j]; e = epi[l, j]
        print(f"{l:3d} P{j} | {m[1]-m[0]:7d} {m[2]-m[1]:7d} | {m[0]-t0:8d} | {e[1]-e[0]:8d} {e[2]-e[1]:7d} {e[2]-t0:8d}")
print("group cycles (MMA issue span):", (mma[9, 1, 2] - t0).item())
base = pr[196].item()   # P0 layer-2 K-loop committed (ns)
print("globaltimer ns relative to 'P0 L2 K-loop committed':")
print("  P1 L2 committed      ", pr[197].item() - base)
print("  aready[0] for L3 seen", pr[194].item() - base, "  aready[1] for L3 seen", pr[195].item() - base)
print("  aready[0] for L2 seen", pr[192].item() - base, "  aready[1] for L2 seen", pr[193].item() - base)
for rank in range(2):
    acc = [pr[200 + rank * 16 + w].item() - base for w in range(16)]
    done = [pr[160 + rank * 16 + w].item() - base for w in range(16)]
    print(f"  rank {rank} acc-seen tile0 warps:", acc[:8], " tile1:", acc[8:])
    print(f"  rank {rank} epi-done tile0 warps:", done[:8], " tile1:", done[8:])
print("layer 3 / P0 weight waits (leader clock64, relative to the MMA thread's aready stamp of layer 3 P0):")
ref = mma[3, 0, 1].item()
for i in range(4):
    w = pr[240 + i * 4: 240 + i * 4 + 3]
    print(f"  chunk {i}: TMA issued at {pr[256 + i].item() - ref:7d} | wait begins {w[0].item() - ref:7d}  full(local) {w[1].item() - ref:7d}  pfull(peer) {w[2].item() - ref:7d}")
