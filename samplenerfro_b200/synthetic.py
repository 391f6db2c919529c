"""Synthetic scenes for benchmarks and smoke tests (setup code, not part of the timed path): an analytic
refractive blob voxelised like voxelize_mesh.py:72-106 (ss^3 supersampled occupancy in [1, 1.33]), rescaled
like train.py:220-225, and Blender-style cameras on a sphere (rnerf/datasets.py:216-242)."""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np
import torch

from .utils import Rays

_RI_033_KEYS = ("glass", "wineglass", "pen", "torus_skydome-bkgd_cycles", "dolphin", "lighthouse", "yellow")


def ior_scale_for_config(cfg_name: str) -> float:
    """train.py:220: the scene name picks the refractive-index scale."""
    return 0.33 if any(k in cfg_name for k in _RI_033_KEYS) else 0.5


def ellipsoid_occupancy(G: int, extent: float, radii: Sequence[float], center=(0.0, 0.0, 0.0), ss: int = 4,
                        device="cuda") -> torch.Tensor:
    """mesh.pkl-style `data` [G^3] in [1, 1.33]: mean inside/outside of ss^3 sub-samples per voxel; x slowest."""
    lin = torch.linspace(-extent, extent, G, device=device, dtype=torch.float64)
    d = float(lin[1] - lin[0])
    offs = ((torch.arange(ss, device=device, dtype=torch.float64) + 0.5) / ss - 0.5) * d
    out = torch.empty(G, G, G, device=device, dtype=torch.float32)
    r = [float(v) for v in radii]
    yy = ((lin[:, None] + offs[None, :] - center[1]) / r[1]) ** 2        # [G, ss]
    zz = ((lin[:, None] + offs[None, :] - center[2]) / r[2]) ** 2
    yz = yy[:, None, :, None] + zz[None, :, None, :]                     # [G, G, ss, ss]
    slab = max(1, min(G, (1 << 25) // (G * G)))                            # ~32 M voxels x ss^3 sub-samples per pass
    for i in range(0, G, slab):
        xx = ((lin[i:i + slab, None] + offs[None, :] - center[0]) / r[0]) ** 2     # [slab, ss]
        inside = (xx[:, None, None, :, None, None] + yz[None, :, :, None, :, :]) < 1.0
        out[i:i + slab] = inside.float().mean(dim=(3, 4, 5))
    return (1.0 + 0.33 * out).reshape(-1)


def rescale_ior(data: torch.Tensor, cfg_name: str) -> torch.Tensor:
    """train.py:223: (data - 1) * ri / 0.33 + 1 (float64 like the reference's numpy), then fp32."""
    ri = ior_scale_for_config(cfg_name)
    return ((data.double() - 1.0) * ri / 0.33 + 1.0).float()


def camera_pose(theta: float, phi: float, radius: float) -> np.ndarray:
    """Camera on a sphere looking at the origin, Blender convention (-z forward, +y up in camera space)."""
    pos = radius * np.array([math.cos(theta) * math.sin(phi), math.sin(theta) * math.sin(phi), math.cos(phi)])
    fwd = -pos / np.linalg.norm(pos)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0])); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, -fwd, pos
    return c2w


def blender_rays(c2w: np.ndarray, h: int, w: int, camera_angle_x: float = 0.6911112, use_pixel_centers=True) -> Rays:
    """Per-pixel rays of one Blender camera (rnerf/datasets.py:216-242) as [H,W,.] float32 numpy-backed tensors."""
    focal = 0.5 * w / math.tan(0.5 * camera_angle_x)
    pc = 0.5 if use_pixel_centers else 0.0
    x, y = np.meshgrid(np.arange(w, dtype=np.float32) + pc, np.arange(h, dtype=np.float32) + pc, indexing="xy")
    cam = np.stack([(x - w * 0.5) / focal, -(y - h * 0.5) / focal, -np.ones_like(x)], axis=-1)
    c2w = np.asarray(c2w, dtype=np.float32)
    directions = (cam[..., None, :] * c2w[None, None, :3, :3]).sum(axis=-1)
    origins = np.broadcast_to(c2w[None, None, :3, -1], directions.shape)
    viewdirs = directions / np.linalg.norm(directions, axis=-1, keepdims=True)
    dx = np.sqrt(np.sum((directions[:-1, :, :] - directions[1:, :, :]) ** 2, -1))
    dx = np.concatenate([dx, dx[-2:-1, :]], 0)
    radii = dx[..., None] * 2 / np.sqrt(12)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return Rays(t(origins), t(directions), t(viewdirs), t(radii))
