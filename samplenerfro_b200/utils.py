"""Host-side mirror of the parts of rnerf/utils.py that sit either side of the hot path: the `Rays`
container, flag defaults + YAML overlay, a gin-subset reader for the unchanged configs/*.gin, the chunked
`render_image` driver, the learning-rate schedule and ray sharding.  (Reference: rnerf/utils.py.)
"""
from __future__ import annotations

import ast
import collections
import dataclasses
import math
import os
import re
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence

import torch
import yaml

Rays = collections.namedtuple("Rays", ("origins", "directions", "viewdirs", "radii"))  # rnerf/utils.py:67


def namedtuple_map(fn, tup):
    """Apply `fn` to each element of `tup` and cast to `tup`'s namedtuple (rnerf/utils.py:70-72)."""
    return type(tup)(*map(fn, tup))


@dataclasses.dataclass
class Config:
    """gin-configurable `Config` of rnerf/utils.py:75-84."""
    kernel_size: int = 3
    kernel_sigma: float = 1.0
    voxel_grid: str = "voxelize"
    radiance_weight_name: Optional[str] = "radiance"
    ior_weight_name: Optional[str] = "ior"
    all_weight_name: Optional[str] = "all"


# flag defaults of rnerf/utils.py:87-245 (define_flags); names unchanged
FLAG_DEFAULTS: Dict[str, Any] = dict(
    gin_file=None, gin_param=None, train_dir=None, stage_dir=None, data_dir=None, config=None,
    dataset="blender", batching="single_image", white_bkgd=True, batch_size=1024, factor=4, spherify=False,
    render_path=False, llffhold=8, use_pixel_centers=False, stage="radiance", skip_frames=1,
    model="nerf", near=2.0, far=6.0, net_depth=8, net_width=256, net_depth_condition=1, net_width_condition=128,
    weight_decay_mult=0.0, skip_layer=4, num_rgb_channels=3, num_sigma_channels=1, randomized=True,
    min_deg_point=0, max_deg_point=10, deg_view=4, num_coarse_samples=64, num_fine_samples=128, use_viewdirs=True,
    sh_deg=-1, sh_direnc_deg=-1, noise_std=None, lindisp=False, net_activation="relu", rgb_activation="sigmoid",
    sigma_activation="softplus", legacy_posenc_order=False,
    lr_init=5e-4, lr_final=5e-6, lr_delay_steps=2500, lr_delay_mult=0.01, grad_max_norm=0.0, grad_max_val=0.0,
    max_steps=1000000, save_every=10000, print_every=100, render_every=5000, gc_every=10000, precrop_iters=0,
    precrop_frac=0.5, num_path_samples=8, sparsity_weight=0.0, use_fine_sparsity=False, use_online_sparsity=True,
    extra_batch_size=1024, normal_loss_weight=0.0, normal_smooth_weight=0.0, anneal_delay_steps=80000,
    anneal_max_steps=160000, beta_weight=0.0, bg_weight=0.0, bg_smooth_weight=0.0, bg_patch_size=0,
    eval_once=True, save_output=True, chunk=8192, eval_train=False,
)


class Flags:
    """Stand-in for absl FLAGS: attribute access over the reference's flag names and defaults."""

    def __init__(self, **overrides):
        self.__dict__.update(FLAG_DEFAULTS)
        unknown = set(overrides) - set(FLAG_DEFAULTS)
        if unknown:
            raise ValueError(f"Unknown flags {sorted(unknown)}")
        self.__dict__.update(overrides)


def update_flags(args: Flags, base_dir: Optional[str] = None) -> None:
    """YAML overlay (rnerf/utils.py:248-257): `<config>.yaml` may only set flags that already exist."""
    pth = args.config + ".yaml" if base_dir is None else os.path.join(base_dir, args.config + ".yaml")
    with open(pth, "r") as fin:
        configs = yaml.load(fin, Loader=yaml.FullLoader)
    invalid_args = list(set(configs.keys()) - set(dir(args)) - set(args.__dict__))
    if invalid_args:
        raise ValueError(f"Invalid args {invalid_args} in {pth}.")
    args.__dict__.update(configs)


_GIN_LINE = re.compile(r"^\s*([A-Za-z_][\w]*)\.([A-Za-z_][\w]*)\s*=\s*(.+?)\s*$")


def parse_gin(files: Optional[Iterable[str]] = None, bindings: Optional[Iterable[str]] = None) -> Dict[str, Dict[str, Any]]:
    """The subset of gin the shipped configs use: `Name.attr = python-literal` lines and `#` comments
    (rnerf/utils.py:267-270 calls gin.parse_config_files_and_bindings; gin-config is not installable here)."""
    out: Dict[str, Dict[str, Any]] = collections.defaultdict(dict)
    lines: List[str] = []
    for f in files or []:
        with open(f, "r") as fh:
            lines += fh.read().splitlines()
    for b in bindings or []:
        lines += b.splitlines()
    for raw in lines:
        line = raw.split("#", 1)[0].strip() if "'" not in raw and '"' not in raw else _strip_comment(raw)
        if not line:
            continue
        m = _GIN_LINE.match(line)
        if not m:
            raise ValueError(f"unsupported gin syntax: {raw!r}")
        scope, attr, val = m.groups()
        out[scope][attr] = ast.literal_eval(val)
    return dict(out)


def _strip_comment(raw: str) -> str:
    quote = None
    for i, ch in enumerate(raw):
        if quote:
            if ch == quote:
                quote = None
        elif ch in "'\"":
            quote = ch
        elif ch == "#":
            return raw[:i].strip()
    return raw.strip()


def load_config(gin_files=None, gin_params=None):
    """-> (Config, gin dict).  Mirrors rnerf/utils.py:267-270."""
    g = parse_gin(gin_files, gin_params)
    known = {f.name for f in dataclasses.fields(Config)}
    bad = set(g.get("Config", {})) - known
    if bad:
        raise ValueError(f"Unknown Config bindings {sorted(bad)}")
    return Config(**g.get("Config", {})), g


def render_image(render_fn: Callable, rays: Rays, rng, normalize_disp: bool, chunk: int = 8192, world_size: int = 1,
                 debug: bool = False):
    """Render all the pixels of an image in `chunk`-ray pieces (rnerf/utils.py:331-389).

    render_fn(key_0, key_1, chunk_rays) -> (ret, loss_sp) with ret[-1] = (rgb, distance, acc, trans, trans_rgb_bkgd).
    Chunks are padded by edge replication to a multiple of `world_size` like the reference pads to the
    device count (rnerf/utils.py:357-361).  Returns (rgb[H,W,3], distance[H,W,1], acc[H,W,1]).

    debug=True is the reference's commented "NOTE(debug): visualize coarse and fine samples" variant
    (rnerf/utils.py:371-389 with rnerf/models.py:362-365,533-534; consumer extract_mesh.py:178): render_fn must then
    return model.apply(..., debug=True)'s 3-tuple, and the result is the 8-tuple
    (rgb, distance, acc, ray_pos[H,W,Nt,3], ray_dir[H,W,Nt,3], idx_grad[H,W,Nt,3], trans[H,W,1], ray_pos_c[H,W,Nc,3]):
    the bent FINE sample positions / directions / grad n (what the fine tuple's ray_pos_c, ray_dir_c, idx_grad_c hold after
    sample_pdf, rnerf/models.py:371) and the coarse sample positions.
    """
    height, width = rays[0].shape[:2]
    num_rays = height * width
    rays = namedtuple_map(lambda r: r.reshape((num_rays, -1)), rays)
    key_0, key_1 = _split_key(rng)
    results = []
    for i in range(0, num_rays, chunk):
        chunk_rays = namedtuple_map(lambda r: r[i:i + chunk], rays)
        chunk_size = chunk_rays[0].shape[0]
        rem = chunk_size % world_size
        padding = world_size - rem if rem != 0 else 0
        if padding:
            chunk_rays = namedtuple_map(lambda r: torch.cat([r, r[-1:].expand(padding, -1)], dim=0), chunk_rays)
        out = render_fn(key_0, key_1, chunk_rays)
        chunk_results = list(out[0][-1])
        if debug:
            if len(out) < 3 or not isinstance(out[2], dict):
                raise ValueError("render_image(debug=True) needs a render_fn that calls model.apply(..., debug=True)")
            dbg = out[2]
            chunk_results += [dbg["pos_f"], dbg["dir_f"], dbg["idx_grad_f"], dbg["ray_pos_c"]]
        results.append([x[:-padding] if padding else x for x in chunk_results])
    cat = [torch.cat(r, dim=0) for r in zip(*results)]
    rgb, distance, acc, trans, trans_rgb_bkgd = cat[:5]
    if normalize_disp:
        distance = (distance - distance.min()) / (distance.max() - distance.min())
    ret = (rgb.reshape(height, width, -1), distance.reshape(height, width, -1), acc.reshape(height, width, -1))
    if debug:
        ray_pos, ray_dir, idx_grad, ray_pos_c = cat[5:9]
        ret += (ray_pos.reshape(height, width, -1, 3), ray_dir.reshape(height, width, -1, 3),
                idx_grad.reshape(height, width, -1, 3), trans.reshape(height, width, 1),
                ray_pos_c.reshape(height, width, -1, 3))
    return ret


def _split_key(rng):
    """jax.random.split(rng, 3)[1:] stand-in: two derived integer seeds."""
    seed = int(rng) if rng is not None else 0
    return (seed * 2654435761 + 1) % (2 ** 31), (seed * 2654435761 + 2) % (2 ** 31)


def compute_psnr(mse):
    """rnerf/utils.py:392-401."""
    if isinstance(mse, torch.Tensor):
        return -10.0 * torch.log(mse) / math.log(10.0)
    return -10.0 * math.log(mse) / math.log(10.0)


def learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1, lr_start_steps=0):
    """Log-linear decay with a sine warm-up (rnerf/utils.py:490-528)."""
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(
            0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
    else:
        delay_rate = 1.0
    start_rate = min(max(step - lr_start_steps, 0), 1)
    t = min(max(max(step - lr_start_steps, 0) / (max_steps - lr_start_steps), 0.0), 1.0)
    log_lerp = math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)
    return start_rate * delay_rate * log_lerp


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous ray range of `rank` (reference: shard() reshapes to [n_dev, B/n_dev, ...], rnerf/utils.py:531-534).
    Like the reference (train.py:196: "Batch size must be divisible by the number of devices") a remainder is an error,
    not silently dropped rays."""
    if n % world_size != 0:
        raise ValueError(f"batch size {n} must be divisible by the number of devices {world_size}")
    per = n // world_size
    return rank * per, (rank + 1) * per


def load_mesh_pkl(path: str):
    """mesh.pkl schema (voxelize_mesh.py:109-116; loader train.py:209-217) -> (data[G^3,1] float64, ndim, nmin, nmax)."""
    import pickle
    with open(path, "rb") as f:
        mesh_dict = pickle.load(f)
    if mesh_dict["extent"] > 0:
        nmin = [-mesh_dict["extent"]] * 3
        nmax = [mesh_dict["extent"]] * 3
    else:
        nmin = list(mesh_dict["min_point"])
        nmax = list(mesh_dict["max_point"])
    ndim = [mesh_dict["num_voxels"]] * 3
    return mesh_dict["data"], ndim, nmin, nmax


def generate_rays(camtoworld, height: int, width: int, focal: Optional[float] = None, cam_mat=None,
                  use_pixel_centers: bool = True, device="cuda") -> Rays:
    """Rays of one camera built on the device: Dataset._generate_rays (rnerf/datasets.py:216-242 / :486-518) without
    the host arrays and their host->device copy.  [H,W,.] fields like the reference's `rays` of one image."""
    from . import ops
    o, d, v, r = ops.generate_rays(camtoworld, height, width, focal=focal, cam_mat=cam_mat,
                                   use_pixel_centers=use_pixel_centers, device=device)
    return Rays(o, d, v, r)


def render_view(render_fn: Callable, camtoworld, height: int, width: int, rng, focal: Optional[float] = None, cam_mat=None,
                use_pixel_centers: bool = True, normalize_disp: bool = False, chunk: int = 8192, device="cuda"):
    """render_image for a camera instead of a ray array: rays are generated on the device, chunk outputs are assembled
    on the device -> (rgb[H,W,3], distance[H,W,1], acc[H,W,1]); only the pose crosses the host/device boundary."""
    rays = generate_rays(camtoworld, height, width, focal=focal, cam_mat=cam_mat, use_pixel_centers=use_pixel_centers,
                         device=device)
    return render_image(render_fn, rays, rng, normalize_disp, chunk=chunk)


def band_rows(height: int, rank: int, world_size: int):
    """Image rows [r0, r1) of `rank` when a frame is cut into `world_size` bands of ceil(H / N) rows (the last bands may be
    short or empty) -> (r0, r1, rows_per_band)."""
    per = -(-height // world_size)
    r0 = min(rank * per, height)
    return r0, min(r0 + per, height), per


def _gather_bands(local: List[torch.Tensor], height: int, width: int, per: int, world_size: int, group=None):
    """all_gather of per-rank row bands ([rows*W, C] each, C = 3 + 1 + 1) into [H*W, C] on every rank; bands are padded to
    `per` rows by edge replication so that all ranks contribute equally sized tiles (like rnerf/utils.py:357-361 pads)."""
    import torch.distributed as dist
    tile = torch.cat(local, dim=-1)              # [rows*W, 5]
    want = per * width
    if tile.shape[0] < want:
        fill = tile[-1:] if tile.shape[0] else tile.new_zeros(1, tile.shape[1])
        tile = torch.cat([tile, fill.expand(want - tile.shape[0], -1)], dim=0)
    tile = tile.contiguous()
    full = torch.empty((world_size * want, tile.shape[1]), device=tile.device, dtype=tile.dtype)
    if tile.is_cuda:       # NCCL writes every band straight into its slot of the frame (no list of tiles + concatenation)
        dist.all_gather_into_tensor(full, tile, group=group)
    else:                  # gloo (CPU tests)
        dist.all_gather(list(full.split(want, dim=0)), tile, group=group)
    return full[:height * width]


def render_image_sharded(render_fn: Callable, rays: Rays, rng, normalize_disp: bool, chunk: int = 8192, rank: int = 0,
                         world_size: int = 1, group=None):
    """render_image with the frame's rows partitioned over `world_size` processes (one per GPU): every rank renders its
    band -- rays are independent, so there is no data-path collective -- and the bands are all-gathered, which is the
    role of `jax.lax.all_gather(..., "batch")` in eval.py:96.  Returns the full (rgb, distance, acc) on every rank."""
    height, width = rays[0].shape[:2]
    r0, r1, per = band_rows(height, rank, world_size)
    if r1 > r0:
        band = namedtuple_map(lambda r: r[r0:r1], rays)
        rgb, dist_, acc = render_image(render_fn, band, rng, False, chunk=chunk)
        local = [rgb.reshape(-1, 3), dist_.reshape(-1, 1), acc.reshape(-1, 1)]
    else:
        ref = rays[0]
        local = [torch.zeros(0, c, device=ref.device, dtype=torch.float32) for c in (3, 1, 1)]
    full = _gather_bands(local, height, width, per, world_size, group) if world_size > 1 else torch.cat(local, dim=-1)
    rgb, distance, acc = full[:, 0:3], full[:, 3:4], full[:, 4:5]
    if normalize_disp:
        distance = (distance - distance.min()) / (distance.max() - distance.min())
    return rgb.reshape(height, width, 3), distance.reshape(height, width, 1), acc.reshape(height, width, 1)


def render_view_sharded(render_fn: Callable, camtoworld, height: int, width: int, rng, focal: Optional[float] = None,
                        cam_mat=None, use_pixel_centers: bool = True, normalize_disp: bool = False, chunk: int = 8192,
                        rank: int = 0, world_size: int = 1, group=None, device="cuda"):
    """render_view over `world_size` GPUs: each rank generates ONLY its band's rays on its device (rnerf_generate_rays
    takes a row range), renders them and the bands are all-gathered (config D of BASELINE.json: one frame, 8 x B200)."""
    from . import ops
    r0, r1, per = band_rows(height, rank, world_size)
    if r1 > r0:
        o, d, v, r = ops.generate_rays(camtoworld, height, width, focal=focal, cam_mat=cam_mat,
                                       use_pixel_centers=use_pixel_centers, row0=r0, n_rows=r1 - r0, device=device)
        rgb, dist_, acc = render_image(render_fn, Rays(o, d, v, r), rng, False, chunk=chunk)
        local = [rgb.reshape(-1, 3), dist_.reshape(-1, 1), acc.reshape(-1, 1)]
    else:
        local = [torch.zeros(0, c, device=device, dtype=torch.float32) for c in (3, 1, 1)]
    full = _gather_bands(local, height, width, per, world_size, group) if world_size > 1 else torch.cat(local, dim=-1)
    rgb, distance, acc = full[:, 0:3], full[:, 3:4], full[:, 4:5]
    if normalize_disp:
        distance = (distance - distance.min()) / (distance.max() - distance.min())
    return rgb.reshape(height, width, 3), distance.reshape(height, width, 1), acc.reshape(height, width, 1)


class GridPoints:
    """datasets.Grid (rnerf/datasets.py:245-328), the producer of batch["pts"] / batch["grads"] for
    compute_normal_loss_and_smooth, on the device: candidate voxels are those with |grad n| > 1e-3 (:264); a batch picks
    `extra_batch_size` of them at random, places the point at idx / ndim * (nmax - nmin) + nmin (the reference divides by
    ndim, not ndim - 1, :272-273) plus uniform(-1, 1) * ndelta (:274) and interpolates grad n there (:275; the same
    central-difference gradient and trilinear interpolation as the model's table, so the table's lookup kernel is used)."""

    def __init__(self, model, extra_batch_size: int = 1024):
        self.model, self.batch_size = model, int(extra_batch_size)
        g = model.table.view(model.ndim[0], model.ndim[1], model.ndim[2], 4)[..., 1:4]
        self.candidate_indices = torch.nonzero(g.norm(dim=-1) > 1e-3)          # [K, 3] (x, y, z) voxel indices
        if self.candidate_indices.shape[0] == 0:
            raise ValueError("the grid has no voxel with |grad n| > 1e-3")

    def next_train(self, generator: Optional[torch.Generator] = None, indices=None, noise=None) -> Dict[str, torch.Tensor]:
        """-> {"pts": [B,1,3], "grads": [B,1,3]}.  `indices` (into candidate_indices) / `noise` ([B,3] in [-1,1)) replace
        the random draws for reproducible parity tests."""
        from . import ops
        m, dev = self.model, self.model.table.device
        if indices is None:
            indices = torch.randint(0, self.candidate_indices.shape[0], (self.batch_size,), generator=generator, device=dev)
        idx = self.candidate_indices[torch.as_tensor(indices, device=dev).long()].double()
        nd, lo, hi = (torch.tensor(v, dtype=torch.float64, device=dev) for v in (m.ndim, m.nmin, m.nmax))
        if noise is None:
            noise = torch.rand(idx.shape[0], 3, generator=generator, device=dev, dtype=torch.float64) * 2 - 1
        ndelta = (hi - lo) / (nd - 1.0)
        pts = (idx / nd * (hi - lo) + lo + torch.as_tensor(noise, device=dev).double() * ndelta).float().contiguous()
        grads = ops.grid_lookup(m.table, m.ndim, m.nmin, m.nmax, pts)[:, 1:4].contiguous()
        return {"pts": pts[:, None], "grads": grads[:, None]}


def image_psnr(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """compute_psnr(mean((pred - target)^2)) with the reduction on the device (eval.py's per-image metric)."""
    from . import ops
    return compute_psnr(ops.image_mse(pred.contiguous(), target.contiguous()))
