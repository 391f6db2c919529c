"""Shared synthetic inputs for the parity tests (seeded; fed identically to the oracle and the CUDA path)."""
import math

import numpy as np
import torch

from oracle import rnerf_oracle as O


def sphere_grid(G=32, extent=1.5, radius=0.8, center=(0.0, 0.0, 0.0), cfg_name="example", ws=3, sigma=1.0, ss=2):
    """Voxelised sphere like voxelize_mesh.py:72-106 (ss^3 supersampled occupancy in [1,1.33]) -> rescale -> blur."""
    lin = np.linspace(-extent, extent, G)
    d = (lin[1] - lin[0])
    occ = np.zeros((G, G, G))
    offs = (np.arange(ss) + 0.5) / ss - 0.5
    X, Y, Z = np.meshgrid(lin, lin, lin, indexing="ij")
    for a in offs:
        for b in offs:
            for c in offs:
                occ += (((X + a * d - center[0]) ** 2 + (Y + b * d - center[1]) ** 2 + (Z + c * d - center[2]) ** 2)
                        < radius ** 2)
    data = 1.0 + 0.33 * occ / ss ** 3
    ndim = [G, G, G]
    nmin = [-extent] * 3
    nmax = [extent] * 3
    n = O.ior_rescale(data.reshape(-1, 1), cfg_name)
    if ws > 0:
        n = O.conv3d_normal(n, ndim, ws, sigma)
    else:
        n = O.as_t(n)
    return n.to(torch.float32).contiguous(), ndim, nmin, nmax


def random_rays(B, seed=0, radius=4.0, target_extent=1.0):
    gen = torch.Generator().manual_seed(seed)
    o = torch.randn(B, 3, generator=gen)
    o = o / o.norm(dim=-1, keepdim=True) * radius
    tgt = (torch.rand(B, 3, generator=gen) * 2 - 1) * target_extent
    d = tgt - o
    d = d / d.norm(dim=-1, keepdim=True)
    return o.contiguous(), d.contiguous()


def camera_rays(h, w, seed=0, radius=4.03):
    """A blender-style camera on a sphere looking at the origin (rnerf/datasets.py:216-242)."""
    rng = np.random.RandomState(seed)
    th, ph = rng.uniform(0, 2 * math.pi), rng.uniform(0.2, 1.2)
    pos = radius * np.array([math.cos(th) * math.sin(ph), math.sin(th) * math.sin(ph), math.cos(ph)])
    fwd = -pos / np.linalg.norm(pos)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up); right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    c2w = np.eye(4)
    c2w[:3, 0] = right; c2w[:3, 1] = upv; c2w[:3, 2] = -fwd; c2w[:3, 3] = pos
    focal = 0.5 * w / math.tan(0.5 * 0.6911112)
    return O.generate_rays(c2w, h, w, focal)


def to_cuda_params(tree):
    if isinstance(tree, dict):
        return {k: to_cuda_params(v) for k, v in tree.items()}
    return tree.detach().to("cuda", torch.float32).contiguous()


def rel_err(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def psnr(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    mse = ((a - b) ** 2).mean().item()
    return 99.0 if mse == 0 else -10.0 * math.log10(mse)
