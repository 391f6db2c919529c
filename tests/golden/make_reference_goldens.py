#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN SOURCE FILES (/root/reference/rnerf/*.py, unmodified)
under the numpy-backed jax/flax shim in tests/golden/_jaxshim.  Run in the build container only:

    python tests/golden/make_reference_goldens.py

writes tests/golden/ref_*.npz.  tests/test_oracle_vs_reference.py then pins the CPU oracle (oracle/rnerf_oracle.py)
against these files.  The shim reproduces jax/flax *semantics* (32-bit types, Flax parameter naming, nn.scan,
.at[].set, lax.fori_loop, the positional-argument quirk of jnp.nan_to_num); random draws are recorded and stored so
the oracle receives the same jitter / u / noise.
"""
import dataclasses
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_jaxshim"))
sys.path.insert(0, "/root/reference")
np.seterr(all="ignore")

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
from rnerf import eikonal_utils, ior_utils, math_utils, model_utils, models, utils  # noqa: E402


def f32(x):
    return np.asarray(x, dtype=np.float32)


def flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(flatten(v, prefix + k + "/"))
        else:
            out[prefix + k] = np.asarray(v)
    return out


def sphere_grid(G, extent, radius, center=(0, 0, 0)):
    lin = np.linspace(-extent, extent, G)
    X, Y, Z = np.meshgrid(lin, lin, lin, indexing="ij")
    r2 = (X - center[0]) ** 2 + (Y - center[1]) ** 2 + (Z - center[2]) ** 2
    return np.where(r2 < radius ** 2, 1.33, 1.0).reshape(-1, 1)


def rays_towards_box(B, seed, radius=4.0, extent=0.9, shift=(0, 0, 0)):
    rs = np.random.RandomState(seed)
    o = rs.normal(size=(B, 3)); o = o / np.linalg.norm(o, axis=-1, keepdims=True) * radius + np.array(shift)
    tgt = rs.uniform(-extent, extent, size=(B, 3)) + np.array(shift)
    d = tgt - o; d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    return f32(o), f32(d)


def args_for(config, **kw):
    a = types.SimpleNamespace(
        net_activation="relu", rgb_activation="sigmoid", sigma_activation="softplus", sh_deg=-1, use_viewdirs=True,
        num_rgb_channels=3, num_sigma_channels=1, min_deg_point=0, max_deg_point=10, deg_view=4, num_coarse_samples=64,
        num_fine_samples=128, near=2.0, far=6.0, noise_std=None, white_bkgd=False, net_depth=8, net_width=256,
        net_depth_condition=1, net_width_condition=128, skip_layer=4, lindisp=False, legacy_posenc_order=False,
        stage="radiance", num_path_samples=12, use_fine_sparsity=False, use_online_sparsity=False, sh_direnc_deg=-1,
        config=config, randomized=False)
    a.__dict__.update(kw)
    return a


def run_model(name, config, G, extent, radius, center, ws, sigma, B, seed, near, far, P, ri_scale, bd_cut_dist=None,
              randomized=False, shift=(0, 0, 0), bias_scale=0.1, keep_path=16):
    ndim, nmin, nmax = [G] * 3, [-extent] * 3, [extent] * 3
    data = sphere_grid(G, extent, radius, center)
    grid = ior_utils.conv3d_normal((data - 1.0) * ri_scale / 0.33 + 1.0, ndim, ws, sigma)       # train.py:223
    o, d = rays_towards_box(B, seed, shift=shift)
    rays = utils.Rays(origins=jnp.array(o), directions=jnp.array(d), viewdirs=jnp.array(d), radii=jnp.array(np.ones((B, 1))))
    args = args_for(config, near=near, far=far, num_path_samples=P, randomized=randomized)
    if bd_cut_dist is not None:
        # gin would bind NerfModel.bd_cut_dist; without gin the class default is patched for this run
        models.NerfModel.bd_cut_dist = bd_cut_dist
        models.NerfModel.__dataclass_fields__["bd_cut_dist"].default = bd_cut_dist
    example = {"rays": utils.namedtuple_map(lambda x: x[None], rays)}
    import flax.linen as _nn
    _nn._CTX["key"] = 0          # same parameter draw for every scene: the 5 MB weight file is stored once
    model, variables = models.construct_nerf(jax.random.PRNGKey(seed), example, args, ndim=ndim, nmin=nmin, nmax=nmax,
                                             grid=grid)
    if bd_cut_dist is not None:
        models.NerfModel.bd_cut_dist = None
        models.NerfModel.__dataclass_fields__["bd_cut_dist"].default = None
        if model.bd_cut_dist != bd_cut_dist:
            object.__setattr__(model, "bd_cut_dist", bd_cut_dist)
    # non-zero biases so every term is exercised (Flax initialises them to zero)
    rs = np.random.RandomState(1234)
    for mlp in ("coarse_mlp", "fine_mlp", "bkgd_mlp"):
        for dn in variables["params"][mlp].values():
            dn["bias"] = jnp.array(rs.uniform(-bias_scale, bias_scale, size=dn["bias"].shape))
    jax.random.DRAWS.clear()
    ret, loss_sp = model.apply(variables, jax.random.PRNGKey(11), jax.random.PRNGKey(12), rays, randomized)
    draws = list(jax.random.DRAWS)
    jitter_off = [v for k, v in draws if k == "randint"][0]
    out = {"grid": f32(grid), "ndim": np.array(ndim), "nmin": np.array(nmin), "nmax": np.array(nmax),
           "origins": o, "viewdirs": d, "near": near, "far": far, "num_path_samples": P, "config": config,
           "jitter": np.arange(0, 64 * P, P) + np.asarray(jitter_off), "loss_sp": f32(loss_sp),
           "bd_cut_dist": -1.0 if bd_cut_dist is None else bd_cut_dist}
    if randomized:
        un = [v for k, v in draws if k == "uniform"][0]
        out["u_noise"] = f32(un)
    for lvl, tup in enumerate(ret):
        for nm, v in zip(("rgb", "distance", "acc", "trans", "trans_rgb_bkgd"), tup):
            out[f"ret{lvl}_{nm}"] = f32(v)
    # the bent path itself (PathSampler), through the model's own parameters
    ps = eikonal_utils.PathSampler(near=near, far=far, stage="radiance", num_samples=64 * P,
                                   step_size=(far - near) / (64 * P - 1), ndim=ndim, nmin=nmin, nmax=nmax, grid=grid)
    pos, dirs, dist, idn, idg = ps.apply({"params": variables["params"]["path_sampler"]}, rays.origins, rays.viewdirs, 1.0)
    # the whole path of the first `keep_path` rays, and a SHA-256 digest of every array over ALL rays (the oracle and the
    # CUDA march are bit-exact against the reference's scan, so digests can be compared; 1024 rays x 768 steps x 11 floats
    # would be 35 MB per scene)
    import hashlib
    kp = keep_path
    out.update(path_pos=f32(pos)[:kp], path_dir=f32(dirs)[:kp], path_dist=f32(dist)[:kp], path_n=f32(idn)[:kp], path_grad=f32(idg)[:kp])
    for nm, arr in (("pos", pos), ("dir", dirs), ("dist", dist), ("n", idn), ("grad", idg)):
        out[f"path_{nm}_sha256"] = hashlib.sha256(np.ascontiguousarray(f32(arr)).tobytes()).hexdigest()
        out[f"path_{nm}_sum64"] = np.float64(np.asarray(arr, dtype=np.float64).sum())
    params = {k: f32(v) for k, v in flatten(variables["params"]).items()}
    ppath = os.path.join(HERE, "ref_params.npz")
    if os.path.exists(ppath):
        old = np.load(ppath)
        assert all(np.array_equal(old[k], params[k]) for k in params), "parameter draw changed between scenes"
    else:
        np.savez_compressed(ppath, **params)
    np.savez_compressed(os.path.join(HERE, f"ref_model_{name}.npz"), **out)
    print(name, "rgb fine mean", out["ret1_rgb"].mean(), "bend", np.abs(f32(dirs)[:, -1] - d).max())


def run_functions():
    rs = np.random.RandomState(0)
    out = {}
    # encodings (rnerf/model_utils.py:187-245)
    x = f32(rs.uniform(-3, 3, size=(5, 7, 3)))
    out["enc_x"] = x
    out["pos_enc_10"] = f32(model_utils.pos_enc(jnp.array(x), 0, 10))
    out["pos_enc_4"] = f32(model_utils.pos_enc(jnp.array(x), 0, 4))
    out["pos_enc_4_legacy"] = f32(model_utils.pos_enc(jnp.array(x), 0, 4, True))
    out["annealed_enc_alpha3p5"] = f32(model_utils.annealed_pos_enc(jnp.array(x), 0, 10, 3.5))
    # compositing (rnerf/model_utils.py:247-309)
    B, Ns = 9, 40
    rgb = f32(rs.uniform(size=(B, Ns, 3))); sig = f32(rs.uniform(size=(B, Ns, 1)) * 6); sig[0] = 0
    t = f32(2 + np.sort(rs.uniform(size=(B, Ns)) * 4, axis=-1)); dirs = f32(rs.normal(size=(B, Ns, 3)))
    bk = f32(rs.uniform(size=(B, 3))); mask = f32(rs.uniform(size=(B, Ns)) > 0.4)
    out.update(vr_rgb=rgb, vr_sigma=sig, vr_t=t, vr_dirs=dirs, vr_bkgd=bk, vr_mask=mask)
    for tag, kw in (("a", dict(white_bkgd=False, rgb_bkgd=jnp.array(bk))), ("b", dict(white_bkgd=True, rgb_bkgd=None)),
                    ("c", dict(white_bkgd=False, rgb_bkgd=jnp.array(bk), mask_bbox=jnp.array(mask)))):
        r = model_utils.volumetric_rendering(jnp.array(rgb), jnp.array(sig), jnp.array(t), jnp.array(dirs), **kw)
        for nm, v in zip(("comp", "dist", "acc", "w", "alpha", "trans", "trb"), r):
            out[f"vr_{tag}_{nm}"] = f32(v)
    # pdf sampling (rnerf/model_utils.py:312-374)
    B, Nb, N = 6, 63, 128
    bins = f32(2 + np.sort(rs.uniform(size=(B, Nb)) * 4, axis=-1)); w = f32(rs.uniform(size=(B, Nb - 1)) ** 3 + 0.01)
    w[0] = 0.0
    out.update(pdf_bins=bins, pdf_w=w)
    out["pdf_det"] = f32(model_utils.sorted_piecewise_constant_pdf(None, jnp.array(bins), jnp.array(w), N, False))
    jax.random.DRAWS.clear()
    out["pdf_rand"] = f32(model_utils.sorted_piecewise_constant_pdf(jax.random.PRNGKey(3), jnp.array(bins), jnp.array(w), N, True))
    out["pdf_rand_noise"] = f32(jax.random.DRAWS[0][1])
    # grid ops (rnerf/ior_utils.py:165-223, 327-363)
    G = 10
    ndim, nmin, nmax = [G, G, G], [-1.5, -1.0, -0.5], [1.5, 2.0, 0.75]
    g0 = f32(1 + 0.5 * rs.uniform(size=(G ** 3, 1)))
    out["grid_in"] = g0
    out.update(grid_ndim=np.array(ndim), grid_nmin=np.array(nmin), grid_nmax=np.array(nmax))
    for ws, s in ((3, 1.0), (5, 3.0)):
        out[f"blur_{ws}"] = f32(ior_utils.conv3d_normal(g0, ndim, ws, s))
    vm = ior_utils.VoxMLP(ndim=ndim, nmin=nmin, nmax=nmax, grid=jnp.array(g0))
    vars_ = vm.init(jax.random.PRNGKey(0), jnp.array(f32(rs.uniform(-1, 1, size=(4, 3)))))
    pts = f32(rs.uniform(-2.0, 2.5, size=(300, 3)))
    out["lookup_pts"] = pts
    out["grad_table"] = f32(vm.apply(vars_, method=vm._compute_grad))
    out["lookup"] = f32(vm.apply(vars_, jnp.array(pts), method=vm._linear3))
    # "all"-stage march: so3 MLP rotation of grad n inside every step (rnerf/eikonal_utils.py:34-39, ior_utils.py:282-312)
    G = 16
    ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
    grid = ior_utils.conv3d_normal((sphere_grid(G, 1.5, 0.8) - 1.0) * 0.5 / 0.33 + 1.0, ndim, 3, 1.0)
    o, d = rays_towards_box(12, 5)
    S = 96
    ps = eikonal_utils.PathSampler(near=2.0, far=6.0, stage="all", num_samples=S, step_size=4.0 / (S - 1), ndim=ndim,
                                   nmin=nmin, nmax=nmax, grid=grid)
    v = ps.init(jax.random.PRNGKey(1), jnp.array(o), jnp.array(d), 0.7)
    so3 = v["params"]["scan"]["idx_model"]["so3_mlp"]
    so3["Dense_4"]["kernel"] = jnp.array(f32(rs.normal(size=so3["Dense_4"]["kernel"].shape) * 0.3))   # visible rotation
    pos, dirs, dist, idn, idg = ps.apply(v, jnp.array(o), jnp.array(d), 0.7)
    out.update(all_grid=f32(grid), all_o=o, all_d=d, all_pos=f32(pos), all_dir=f32(dirs), all_dist=f32(dist), all_grad=f32(idg))
    for k, val in flatten(so3).items():
        out["all_so3:" + k] = f32(val)
    # Directional derivatives of a linear functional of the coarse samples, L = sum(ray_pos[:, jit] * gp + ray_dir[:, jit] * gd),
    # with respect to so3_mlp, by central differences THROUGH THE REFERENCE'S OWN SCAN (the shim has no autodiff): pins the
    # gradients of the "all"-stage training path (oracle autograd, CUDA reverse sweep) to the reference's code.  The scan is
    # piecewise smooth in the parameters (ReLU, trilinear cells), so the differences are taken in float64 (the shim's
    # jax_enable_x64 stand-in) with a small step; inputs and weights are the float32 values above.
    jit = np.array([0, 9, 17, 30, 41, 55, 70, 95])
    gp, gd = f32(rs.normal(size=(12, 8, 3))), f32(rs.normal(size=(12, 8, 3)))
    base = {k: {kk: np.asarray(f32(vv), np.float64) for kk, vv in lay.items()} for k, lay in so3.items()}
    jnp.X64 = True
    try:
        ps64 = eikonal_utils.PathSampler(near=2.0, far=6.0, stage="all", num_samples=S, step_size=4.0 / (S - 1), ndim=ndim,
                                         nmin=nmin, nmax=nmax, grid=jnp.array(np.asarray(f32(grid), np.float64)))
        o64, d64 = jnp.array(np.asarray(o, np.float64)), jnp.array(np.asarray(d, np.float64))

        def functional(params):
            v64 = {"params": {"scan": {"idx_model": {"so3_mlp": {k: {kk: jnp.array(vv) for kk, vv in lay.items()}
                                                                    for k, lay in params.items()}}}}}
            p_, d_, _, _, _ = ps64.apply(v64, o64, d64, 0.7)
            assert np.asarray(p_).dtype == np.float64
            return float((np.asarray(p_)[:, jit] * gp).sum() + (np.asarray(d_)[:, jit] * gd).sum())

        out.update(fd_jitter=jit, fd_gp=gp, fd_gd=gd)
        for i in range(3):
            delta = {k: {kk: np.asarray(f32(rs.normal(size=vv.shape) * max(float(np.sqrt((vv ** 2).mean())), 0.05)), np.float64)
                         for kk, vv in lay.items()} for k, lay in base.items()}
            vals = []
            for eps in (2e-6, 1e-6):
                plus = functional({k: {kk: base[k][kk] + eps * delta[k][kk] for kk in lay} for k, lay in base.items()})
                minus = functional({k: {kk: base[k][kk] - eps * delta[k][kk] for kk in lay} for k, lay in base.items()})
                vals.append((plus - minus) / (2 * eps))
            print(f"  fd direction {i}: dL = {vals[0]:.8f} (eps 2e-6), {vals[1]:.8f} (eps 1e-6)")
            assert abs(vals[0] - vals[1]) < 5e-3 * abs(vals[1]), "finite differences not converged"
            out[f"fd_dL_{i}"] = np.float64(vals[1])
            for k, lay in delta.items():
                for kk, vv in lay.items():
                    out[f"fd_delta_{i}:{k}/{kk}"] = f32(vv)
    finally:
        jnp.X64 = False
    # math helpers (rnerf/math_utils.py:6-20)
    z = f32(rs.normal(size=(20, 3))); z[0] = 0
    out["mh_x"] = z
    out["mh_normalize"] = f32(math_utils.safe_l2_normalize(jnp.array(z)))
    out["mh_log"] = f32(math_utils.safe_log(jnp.array(np.abs(z))))
    # mip helpers for curved rays (rnerf/mip.py:26-57,60-113,116-175; dead on the live path): the call the
    # commented lines rnerf/models.py:249-254 would make -- t_vals = [ray_dist_c, last + 1e-3], bent positions and per-sample
    # directions, cone and cylinder, diag=True -- then integrated_pos_enc(samples, 0, 10)
    from rnerf import mip
    B, N = 7, 24
    tv = f32(2 + np.sort(rs.uniform(size=(B, N)) * 4, axis=-1))
    tv = np.concatenate([tv, tv[:, -1:] + f32(1e-3)], -1)
    mo = f32(rs.normal(size=(B, N, 3))); md = f32(rs.normal(size=(B, N, 3))); md /= np.linalg.norm(md, axis=-1, keepdims=True)
    md = f32(md * rs.uniform(0.9, 1.1, size=(B, N, 1)))
    mr = f32(rs.uniform(5e-4, 3e-3, size=(B, 1)))
    out.update(mip_t_vals=tv, mip_origins=mo, mip_dirs=md, mip_radii=mr, mip_near=np.float32(2.0))
    for shape in ("cone", "cylinder"):
        mean, cov = mip.cast_rays(jnp.array(tv), jnp.array(mo), jnp.array(md), jnp.array(mr), shape, 2.0)
        out[f"mip_{shape}_mean"], out[f"mip_{shape}_cov"] = f32(mean), f32(cov)
        out[f"mip_{shape}_ipe"] = f32(mip.integrated_pos_enc((mean, cov), 0, 10))
    ex = f32(rs.uniform(-400, 400, size=(50,))); ev = f32(rs.uniform(0, 4, size=(50,)) ** 2)
    ey, eyv = mip.expected_sin(jnp.array(ex), jnp.array(ev))
    out.update(mip_es_x=ex, mip_es_var=ev, mip_es_y=f32(ey), mip_es_yvar=f32(eyv))
    np.savez_compressed(os.path.join(HERE, "ref_functions.npz"), **out)
    print("functions ok")


def run_config_a():
    """BASELINE.json configs[0] at its stated size: configs/example.{gin,yaml} (read with this repo's
    own gin-subset / YAML loaders, so the fixture also pins them), the single camera of example_data/transforms_train.json
    rendered at 100x100 with use_pixel_centers, G = 128, extent 1.5, IoR scale 0.5, blur 3/1; grid = a radius-0.5 sphere
    voxelised with 4^3 supersampling like voxelize_mesh.py:72-106 (the real mesh.pkl is a missing blob).
    All 10 000 rays go through the reference's NerfModel.__call__.  -> ref_model_config_a.npz + config_a.json"""
    import json
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from samplenerfro_b200 import utils as U
    cfg, gin = U.load_config(["/root/reference/configs/example.gin"])
    flags = U.Flags(config="/root/reference/configs/example")
    U.update_flags(flags)
    meta = json.load(open("/root/reference/example_data/transforms_train.json"))
    c2w = np.array(meta["frames"][0]["transform_matrix"], dtype=np.float32)
    Hh = Ww = 100
    # rnerf/datasets.py:355-356.  A Python float on purpose: the reference's pinned NumPy 1.x keeps `float32 array /
    # np.float64 scalar` in float32 (value-based casting), NumPy 2 would promote the whole ray computation to float64
    focal = float(.5 * Ww / np.tan(.5 * float(meta["camera_angle_x"])))
    pc = 0.5 if flags.use_pixel_centers else 0.0
    x, y = np.meshgrid(np.arange(Ww, dtype=np.float32) + pc, np.arange(Hh, dtype=np.float32) + pc, indexing="xy")
    cam = np.stack([(x - Ww * 0.5) / focal, -(y - Hh * 0.5) / focal, -np.ones_like(x)], axis=-1)  # rnerf/datasets.py:218-230
    dirs = (cam[..., None, :] * c2w[None, None, :3, :3]).sum(axis=-1)
    orig = np.broadcast_to(c2w[None, None, :3, -1], dirs.shape)
    view = dirs / np.linalg.norm(dirs, axis=-1, keepdims=True)
    o, dr, v = f32(orig.reshape(-1, 3)), f32(dirs.reshape(-1, 3)), f32(view.reshape(-1, 3))
    G, extent, ss = 128, 1.5, 4
    lin = np.linspace(-extent, extent, G)
    dl = lin[1] - lin[0]
    offs = (np.arange(ss) + 0.5) / ss - 0.5
    occ = np.zeros((G, G, G))
    X, Y, Z = np.meshgrid(lin, lin, lin, indexing="ij")
    for a_ in offs:
        for b_ in offs:
            for c_ in offs:
                occ += ((X + a_ * dl) ** 2 + (Y + b_ * dl) ** 2 + (Z + c_ * dl) ** 2) < 0.5 ** 2
    data = (1.0 + 0.33 * occ / ss ** 3).reshape(-1, 1)
    ndim, nmin, nmax = [G] * 3, [-extent] * 3, [extent] * 3
    grid = ior_utils.conv3d_normal((data - 1.0) * 0.5 / 0.33 + 1.0, ndim, cfg.kernel_size, cfg.kernel_sigma)
    rays = utils.Rays(origins=jnp.array(o), directions=jnp.array(dr), viewdirs=jnp.array(v), radii=jnp.array(np.ones((Hh * Ww, 1))))
    args = args_for("example", near=flags.near, far=flags.far, num_path_samples=flags.num_path_samples, randomized=False,
                    num_coarse_samples=flags.num_coarse_samples, num_fine_samples=flags.num_fine_samples, white_bkgd=flags.white_bkgd)
    import flax.linen as _nn
    _nn._CTX["key"] = 0
    model, variables = models.construct_nerf(jax.random.PRNGKey(0), {"rays": utils.namedtuple_map(lambda t: t[None, :8], rays)},
                                             args, ndim=ndim, nmin=nmin, nmax=nmax, grid=grid)
    assert model.use_mask_bbox is False
    rs = np.random.RandomState(1234)
    for mlp in ("coarse_mlp", "fine_mlp", "bkgd_mlp"):
        for dn in variables["params"][mlp].values():
            dn["bias"] = jnp.array(rs.uniform(-0.1, 0.1, size=dn["bias"].shape))
    params = {k: f32(val) for k, val in flatten(variables["params"]).items()}
    old = np.load(os.path.join(HERE, "ref_params.npz"))
    assert all(np.array_equal(old[k], params[k]) for k in params)
    jax.random.DRAWS.clear()
    ret, loss_sp = model.apply(variables, jax.random.PRNGKey(11), jax.random.PRNGKey(12), rays, False)   # eval.py:97
    P = flags.num_path_samples
    out = {"grid": f32(grid), "jitter": np.arange(0, 64 * P, P) + np.asarray([val for k, val in jax.random.DRAWS if k == "randint"][0]),
           "loss_sp": f32(loss_sp)}
    for lvl, tup in enumerate(ret):
        for nm, val in zip(("rgb", "distance", "acc", "trans", "trans_rgb_bkgd"), tup):
            out[f"ret{lvl}_{nm}"] = f32(val)
    ps = eikonal_utils.PathSampler(near=flags.near, far=flags.far, stage="radiance", num_samples=64 * P,
                                   step_size=(flags.far - flags.near) / (64 * P - 1), ndim=ndim, nmin=nmin, nmax=nmax, grid=grid)
    pos, pdirs, dist, idn, idg = ps.apply({"params": variables["params"]["path_sampler"]}, rays.origins, rays.viewdirs, 1.0)
    import hashlib
    for nm, arr in (("pos", pos), ("dir", pdirs), ("dist", dist), ("n", idn), ("grad", idg)):
        out[f"path_{nm}_sha256"] = hashlib.sha256(np.ascontiguousarray(f32(arr)).tobytes()).hexdigest()
    out["path_pos_last"], out["path_dir_last"] = f32(pos)[:, -1], f32(pdirs)[:, -1]
    np.savez_compressed(os.path.join(HERE, "ref_model_config_a.npz"), **out)
    fx = {"source": "configs/example.gin + configs/example.yaml + example_data/transforms_train.json of the reference, parsed by "
                    "samplenerfro_b200.utils.load_config / update_flags (tests/golden/make_reference_goldens.py:run_config_a)",
          "flags": {k: val for k, val in flags.__dict__.items() if k != "config"}, "gin": gin,
          "config": dataclasses.asdict(cfg), "camera_angle_x": float(meta["camera_angle_x"]),
          "camtoworld": np.asarray(meta["frames"][0]["transform_matrix"], dtype=np.float64).tolist(), "height": Hh, "width": Ww,
          "grid": {"G": G, "extent": extent, "sphere_radius": 0.5, "supersampling": ss, "ior_scale": 0.5}}
    with open(os.path.join(HERE, "config_a.json"), "w") as f:
        json.dump(fx, f, indent=1, sort_keys=True)
    bend = np.abs(f32(pdirs)[:, -1] - v).max()
    print("config A: 10000 rays, rgb fine mean", out["ret1_rgb"].mean(), "rays bent:", int((np.abs(f32(pdirs)[:, -1] - v).max(-1) > 1e-3).sum()),
          "max bend", bend)


def run_train_loss():
    """The forward of train.py's loss (train.py:75-162), by executing the UNMODIFIED source text of `train_step`
    (train.py:58-183) -- cut out of the file with `ast`, because importing train.py would run absl flag parsing and pull
    in optax / tensorboard -- in a namespace that supplies FLAGS and the shim.  The shim has no autodiff: jax.value_and_grad
    evaluates the function and returns zero gradients, so what is pinned is the VALUE of the loss and every entry of
    `stats` (the gradients are pinned by finite differences elsewhere).  -> ref_train_loss.npz"""
    import ast
    src = open("/root/reference/train.py").read()
    node = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "train_step"][0]
    fn_src = "\n".join(src.splitlines()[node.lineno - 1:node.end_lineno])
    captured = {}

    def value_and_grad(f, has_aux=False):
        def run(params):
            val, aux = f(params)
            captured["total"] = val
            return (val, aux), jax.tree_util.tree_map(lambda z: np.zeros_like(np.asarray(z)), params)
        return run

    def tree_reduce(f, tree, initializer=0):
        acc = initializer
        stack = [tree]
        while stack:
            t = stack.pop(0)
            if isinstance(t, dict):
                stack = list(t.values()) + stack
            else:
                acc = f(acc, t)
        return acc

    jax.value_and_grad = value_and_grad
    jax.tree_util.tree_reduce = staticmethod(tree_reduce)
    jax.lax.pmean = lambda x, axis_name=None: x
    G, extent, B, P = 20, 1.5, 64, 12
    ndim, nmin, nmax = [G] * 3, [-extent] * 3, [extent] * 3
    grid = ior_utils.conv3d_normal((sphere_grid(G, extent, 0.8) - 1.0) * 0.5 / 0.33 + 1.0, ndim, 3, 1.0)
    o, d = rays_towards_box(B, 21)
    rs = np.random.RandomState(77)
    rays = utils.Rays(origins=jnp.array(o), directions=jnp.array(d), viewdirs=jnp.array(d), radii=jnp.array(np.ones((B, 1))))
    env = f32(rs.normal(size=(8, 8, 3))); env /= np.linalg.norm(env, axis=-1, keepdims=True)
    env_rays = utils.Rays(origins=jnp.array(env), directions=jnp.array(env), viewdirs=jnp.array(env), radii=jnp.array(env[..., :1]))
    pixels = f32(rs.uniform(size=(B, 3)))
    out = {"grid": f32(grid), "ndim": np.array(ndim), "nmin": np.array(nmin), "nmax": np.array(nmax), "origins": o,
           "viewdirs": d, "pixels": pixels, "env_viewdirs": f32(env), "num_path_samples": P}
    for case, (alpha, bgw, smw, wd) in {"a": (0.5, 0.025, 1.0, 0.0), "b": (0.0, 0.025, 1.0, 0.0), "c": (0.7, 1.0, 0.5, 0.1)}.items():
        args = args_for("example", randomized=True, num_path_samples=P)
        flags = types.SimpleNamespace(stage="radiance", randomized=True, bg_weight=bgw, beta_weight=0.0, use_online_sparsity=False,
                                      sparsity_weight=0.0, normal_loss_weight=0.0, normal_smooth_weight=0.0,
                                      bg_smooth_weight=smw, weight_decay_mult=wd, grad_max_val=0.0, grad_max_norm=0.0)
        import flax.linen as _nn
        _nn._CTX["key"] = 0
        model, variables = models.construct_nerf(jax.random.PRNGKey(3), {"rays": utils.namedtuple_map(lambda x: x[None], rays)},
                                                 args, ndim=ndim, nmin=nmin, nmax=nmax, grid=grid)
        brs = np.random.RandomState(1234)
        for mlp in ("coarse_mlp", "fine_mlp", "bkgd_mlp"):
            for dn in variables["params"][mlp].values():
                dn["bias"] = jnp.array(brs.uniform(-0.1, 0.1, size=dn["bias"].shape))
        params = {k: f32(v) for k, v in flatten(variables["params"]).items()}
        old = np.load(os.path.join(HERE, "ref_params.npz"))
        assert all(np.array_equal(old[k], params[k]) for k in params), "parameter draw differs from ref_params.npz"
        # random-init sigma composites to trans ~ 0.3 everywhere, which would leave mask_bg = trans > 0.5 empty: lower the fine
        # sigma head's bias (recorded) so that the background term sees a mix of rays
        sig_bias = {"a": -2.2, "b": -2.2, "c": -1.6}[case]
        variables["params"]["fine_mlp"]["Dense_8"]["bias"] = jnp.array(f32([sig_bias]))
        out[f"{case}_fine_sigma_bias"] = np.float32(sig_bias)
        state = types.SimpleNamespace(params=variables, apply_gradients=lambda grads: "new_state")
        batch = {"rays": rays, "pixels": jnp.array(pixels), "env_rays": env_rays, "annealed_alpha": jnp.array(f32([alpha]))}
        ns = {"FLAGS": flags, "jax": jax, "jnp": jnp, "random": jax.random, "utils": utils, "math_utils": math_utils}
        exec(compile(fn_src, "/root/reference/train.py", "exec"), ns)
        jax.random.DRAWS.clear()
        new_state, stats, _ = ns["train_step"](model, jax.random.PRNGKey(5), state, batch)
        assert new_state == "new_state"
        draws = list(jax.random.DRAWS)
        out[f"{case}_jitter"] = np.arange(0, 64 * P, P) + np.asarray([v for k, v in draws if k == "randint"][0])
        out[f"{case}_u_noise"] = f32([v for k, v in draws if k == "uniform"][0])
        out[f"{case}_cfg"] = np.array([alpha, bgw, smw, wd], dtype=np.float64)
        out[f"{case}_total"] = f32(captured["total"])
        for k in ("loss", "psnr", "loss_c", "psnr_c", "weight_l2", "loss_sp", "loss_nrm", "annealing_rate", "loss_bg",
                  "loss_bg_c", "loss_bg_smooth"):
            out[f"{case}_{k}"] = f32(getattr(stats, k))
        print("train loss", case, "total", float(out[f"{case}_total"]), "rays with trans > 0.5:", int(captured.get("n_bg", -1)), {k: float(out[f"{case}_{k}"]) for k in ("loss", "loss_c", "loss_bg", "loss_bg_smooth", "weight_l2")})
    np.savez_compressed(os.path.join(HERE, "ref_train_loss.npz"), **out)


if __name__ == "__main__":
    if os.path.exists(os.path.join(HERE, "ref_params.npz")):
        os.remove(os.path.join(HERE, "ref_params.npz"))
    run_functions()
    # config A shape (example.gin/yaml): near/far 2/6, P=12, blur 3/1, IoR scale 0.5
    run_model("example", "example", G=20, extent=1.5, radius=0.8, center=(0, 0, 0), ws=3, sigma=1.0, B=1024, seed=3,
              near=2.0, far=6.0, P=12, ri_scale=0.5, keep_path=24)
    # training-mode sampling (randomized=True): stratified u
    run_model("example_rand", "example", G=20, extent=1.5, radius=0.8, center=(0, 0, 0), ws=3, sigma=1.0, B=1024, seed=4,
              near=2.0, far=6.0, P=12, ri_scale=0.5, randomized=True)
    # config D shape (ball.gin/yaml): near/far 0.2/12, P=24, blur 5/3, bd_cut_dist passes with the hard-coded ball box
    run_model("ball", "ball", G=16, extent=2.0, radius=1.0, center=(0, 1.036, 0), ws=5, sigma=3.0, B=1024, seed=5,
              near=0.2, far=12.0, P=24, ri_scale=0.5, bd_cut_dist=6.0, shift=(0, 1.0, 0))
    run_train_loss()
    run_config_a()
