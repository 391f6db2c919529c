"""Per-kernel parity: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import rnerf_oracle as O
import rnerf_test_helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(cuda_lib):
    n, ndim, nmin, nmax = H.sphere_grid(G=40, ws=3, sigma=1.0)
    table = O.build_table(n, ndim, nmin, nmax)
    return {"n": n, "ndim": ndim, "nmin": nmin, "nmax": nmax, "table": table, "table_cu": table.cuda().contiguous()}


def test_grid_table_bit_exact(cuda_lib, scene):
    from samplenerfro_b200 import ops
    t = ops.grid_table(scene["n"].cuda(), scene["ndim"], scene["nmin"], scene["nmax"]).cpu()
    assert torch.equal(t, scene["table"]), f"max abs diff {(t - scene['table']).abs().max()}"


def test_grid_table_constant_division_equals_ieee(cuda_lib, monkeypatch):
    """The (n, grad n) table divides by the constants 2 * ndelta with the exhaustively verified multiply/FMA sequence; it must
    give the bits of the IEEE division (RNERF_MARCH_DIV=ieee) on noisy grids with non-cubic shapes and awkward pitches, and the
    oracle's table."""
    from samplenerfro_b200 import ops
    gen = torch.Generator().manual_seed(3)
    for ndim, lo, hi in (([17, 9, 33], [-1.3, -0.7, -2.1], [1.9, 0.4, 0.3]), ([40, 40, 40], [-1.5] * 3, [1.5] * 3)):
        n = 1.0 + 0.6 * torch.rand(ndim[0] * ndim[1] * ndim[2], generator=gen)
        n[::5] = n[1::5][: n[::5].numel()]                 # exact zeros among the differences
        fast = ops.grid_table(n.cuda(), ndim, lo, hi).cpu()
        monkeypatch.setenv("RNERF_MARCH_DIV", "ieee")
        ieee = ops.grid_table(n.cuda(), ndim, lo, hi).cpu()
        monkeypatch.delenv("RNERF_MARCH_DIV")
        assert torch.equal(fast.view(torch.int32), ieee.view(torch.int32))
        assert torch.equal(fast, O.build_table(n.reshape(ndim), ndim, lo, hi))


@pytest.mark.parametrize("ws,sigma", [(3, 1.0), (5, 3.0), (9, 3.0)])
def test_grid_blur(cuda_lib, ws, sigma):
    from samplenerfro_b200 import ops
    G = 24
    n0, ndim, _, _ = H.sphere_grid(G=G, ws=0)
    ref = O.conv3d_normal(n0, ndim, ws, sigma)
    out = ops.grid_blur(n0.cuda(), ndim, ws, sigma).cpu()
    assert (out - ref).abs().max().item() < 6e-6   # same taps, different fp32 summation order (up to 729 terms)


def test_lookup_bit_exact(cuda_lib, scene):
    from samplenerfro_b200 import ops
    gen = torch.Generator().manual_seed(1)
    pts = (torch.rand(5001, 3, generator=gen) * 2 - 1) * 1.9   # includes points outside the box (clamp-to-edge)
    pts[:8] = torch.tensor(scene["nmin"]) + torch.arange(8)[:, None] * 0.0769231  # voxel corners / edges
    ref = O.linear3(scene["table"], scene["ndim"], scene["nmin"], scene["nmax"], pts)
    out = ops.grid_lookup(scene["table_cu"], scene["ndim"], scene["nmin"], scene["nmax"], pts.cuda()).cpu()
    assert torch.equal(out, ref), f"max abs diff {(out - ref).abs().max()}"


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("B,S,near,far", [(1000, 96, 2.0, 6.0), (257, 768, 2.0, 6.0), (33, 1536, 0.2, 12.0), (1, 7, 2.0, 6.0),
                                          (70, 770, 2.0, 6.0)])
def test_march_bit_exact(cuda_lib, scene, B, S, near, far, compact):
    """Bent sample positions: north_star asks 1e-4 relative; the kernel reproduces the fp32 oracle bit for bit, with
    full (12-float) and compact (8-float, no idx_grad) records, and the dense ray_dist column equals the records'."""
    from samplenerfro_b200 import ops
    o, d = H.random_rays(B, seed=B)
    pos, dirs, dist, n, g = O.march(scene["table"], scene["ndim"], scene["nmin"], scene["nmax"], o, d, near, far, S)
    path = ops.march(scene["table_cu"], scene["ndim"], scene["nmin"], scene["nmax"], o.cuda(), d.cuda(), near, far, S,
                     compact=compact)
    assert path.rec.shape == (B, S, 8 if compact else 12)
    rp, rd, rt, idn, idg = [None if x is None else x.cpu() for x in ops.path_views(path)]
    bent = (dirs[:, -1] - d).abs().max().item()
    assert B < 100 or bent > 1e-3, "test scene does not bend any ray"
    checks = [("ray_pos", rp, pos), ("ray_dir", rd, dirs), ("ray_dist", rt, dist), ("idx_data", idn, n),
              ("t_col", path.t.cpu(), dist)]
    if not compact:
        checks.append(("idx_grad", idg, g))
    for name, a, b in checks:
        assert torch.equal(a, b), f"{name}: max rel diff {H.rel_err(a, b):.3e}"


def test_march_constant_divisor_division_is_exact(cuda_lib, scene, monkeypatch):
    """The grid coordinates (p - nmin)/ndelta use a 3-instruction division by the (exhaustively verified) constant
    ndelta; forcing IEEE divisions must not change a bit.  Rays starting exactly on / next to nmin exercise the
    zero / tiny-numerator guard."""
    from samplenerfro_b200 import ops
    o, d = H.random_rays(2000, seed=21)
    o[:8] = torch.tensor(scene["nmin"]) - 2.0 * d[:8]       # p0 = o + near*d lands on the grid corner
    o[8:16, 0] = scene["nmin"][0] - 2.0 * d[8:16, 0] + 1e-30
    args = (scene["table_cu"], scene["ndim"], scene["nmin"], scene["nmax"], o.cuda(), d.cuda(), 2.0, 6.0, 768)
    a = ops.march(*args, compact=True)
    monkeypatch.setenv("RNERF_MARCH_DIV", "ieee")
    b = ops.march(*args, compact=True)
    monkeypatch.delenv("RNERF_MARCH_DIV")
    assert torch.equal(a.rec, b.rec) and torch.equal(a.t, b.t)


def test_march_brick_skipping_is_bit_identical(cuda_lib):
    """The brick map only skips gathers: the emitted path must not change by a single bit."""
    from samplenerfro_b200 import ops
    n, ndim, nmin, nmax = H.sphere_grid(G=72, radius=0.6, ws=5, sigma=3.0)
    table = ops.grid_table(n.cuda(), ndim, nmin, nmax)
    bricks = ops.grid_bricks(table, ndim)
    homog = torch.isfinite(bricks).float().mean().item()
    assert 0.3 < homog < 0.99, homog                       # the scene has both kinds of bricks
    vals = bricks[torch.isfinite(bricks)].unique()
    assert vals.numel() >= 2                               # vacuum (n = 1) and the object's core (n = 1.5)
    o, d = H.random_rays(3000, seed=9, target_extent=1.4)
    o[:50] *= 3.0                                          # some rays never enter the grid (clamp-to-edge lookups)
    a = ops.march(table, ndim, nmin, nmax, o.cuda(), d.cuda(), 2.0, 6.0, 768)
    b = ops.march(table, ndim, nmin, nmax, o.cuda(), d.cuda(), 2.0, 6.0, 768, bricks=bricks)
    assert torch.equal(a.rec, b.rec) and torch.equal(a.t, b.t)
    c = ops.march(table, ndim, nmin, nmax, o.cuda(), d.cuda(), 2.0, 6.0, 768, bricks=bricks, compact=True)
    assert torch.equal(c.rec, a.rec[..., :8]) and torch.equal(c.t, a.t)   # compact records = the first 8 floats
    assert (a.rec[:, -1, 4:7].cpu() - d).abs().max() > 1e-2    # and rays do bend


def test_march_constant_grid_kat(cuda_lib):
    """n == n0 -> straight ray, p_k = o + near d + k (step/n0) d, direction unchanged."""
    from samplenerfro_b200 import ops
    G, n0, S = 8, 1.25, 64
    ndim, nmin, nmax = [G] * 3, [-1.0] * 3, [1.0] * 3
    table = ops.grid_table(torch.full((G ** 3,), n0, device="cuda"), ndim, nmin, nmax)
    o, d = H.random_rays(64, seed=3)
    path = ops.march(table, ndim, nmin, nmax, o.cuda(), d.cuda(), 2.0, 6.0, S).rec.cpu().double()
    step = 4.0 / (S - 1)
    k = torch.arange(S, dtype=torch.float64)[None, :, None]
    expect = o.double()[:, None] + 2.0 * d.double()[:, None] + k * (step / n0) * d.double()[:, None]
    assert (path[..., 0:3] - expect).abs().max().item() < 2e-5
    assert (path[..., 3] - (2.0 + k[..., 0] * step / n0)).abs().max().item() < 2e-5
    assert (path[..., 4:7] - d.double()[:, None]).abs().max().item() < 1e-6
    assert (path[..., 7] - n0).abs().max().item() < 1e-6 and path[..., 8:11].abs().max().item() == 0.0


def test_select(cuda_lib, scene):
    from samplenerfro_b200 import ops
    o, d = H.random_rays(100, seed=5)
    path = ops.march(scene["table_cu"], scene["ndim"], scene["nmin"], scene["nmax"], o.cuda(), d.cuda(), 2.0, 6.0, 96)
    jit = (torch.arange(0, 96, 12) + torch.randint(0, 12, (8,), generator=torch.Generator().manual_seed(0))).int()
    pos, dirs, t, grad = ops.select(path, jit.cuda(), want_grad=True)
    pc = path.rec.cpu()
    assert torch.equal(pos.cpu(), pc[:, jit.long(), 0:3]) and torch.equal(dirs.cpu(), ops.path_dirs(path).cpu()[:, jit.long()])
    assert torch.equal(t.cpu(), pc[:, jit.long(), 3]) and torch.equal(grad.cpu(), pc[:, jit.long(), 8:11])
    cpath = ops.march(scene["table_cu"], scene["ndim"], scene["nmin"], scene["nmax"], o.cuda(), d.cuda(), 2.0, 6.0, 96,
                      compact=True)
    pos2, dirs2, t2, _ = ops.select(cpath, jit.cuda())
    assert torch.equal(pos2, pos) and torch.equal(dirs2, dirs) and torch.equal(t2, t)
    with pytest.raises(Exception):
        ops.select(cpath, jit.cuda(), want_grad=True)       # compact records carry no idx_grad


def _composite_inputs(B, Ns, seed):
    gen = torch.Generator().manual_seed(seed)
    raw = torch.randn(B, Ns, 4, generator=gen) * 2
    raw[..., 3] = raw[..., 3] * 3 + 1
    t = 2 + torch.sort(torch.rand(B, Ns, generator=gen) * 4, dim=-1).values
    dirs = torch.randn(B, Ns, 3, generator=gen)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    bk = torch.randn(B, 3, generator=gen)
    return raw, t, dirs, bk


@pytest.mark.parametrize("B,Ns,use_mask,use_bkgd", [(301, 64, False, True), (77, 192, True, True), (5, 50, False, False),
                                                     (16, 1, False, True)])
def test_composite_fwd(cuda_lib, B, Ns, use_mask, use_bkgd):
    from samplenerfro_b200 import ops
    raw, t, dirs, bk = _composite_inputs(B, Ns, B + Ns)
    raw[0, :, 3] = -60.0   # sigma == 0 -> acc == 0 -> distance NaN -> 0 -> clipped to t_0 (T12)
    mask = (torch.rand(B, Ns, generator=torch.Generator().manual_seed(9)) > 0.3).float() if use_mask else None
    ref = O.volumetric_rendering(O.rgb_act(raw[..., :3]), O.sigma_act(raw[..., 3:4]), t, dirs, False,
                                 O.rgb_act(bk) if use_bkgd else None, mask)
    o = ops.composite_fwd(raw.cuda(), t.cuda(), dirs.cuda(), bk.cuda() if use_bkgd else None,
                          mask.cuda() if use_mask else None, want_alpha=True)
    names = ["comp_rgb", "distance", "acc", "weights", "alpha", "trans", "trans_rgb_bkgd"]
    for nm, r in zip(names, ref):
        got = o[nm].cpu()
        assert (got - r).abs().max().item() < 3e-6 * max(1.0, r.abs().max().item()), (nm, (got - r).abs().max().item())
    assert o["distance"][0].item() == t[0, 0].item()
    assert (o["acc"].cpu() + o["trans"].cpu()[:, 0] - 1).abs().max().item() < 1e-5  # acc + T_end == 1 (telescoping)


@pytest.mark.parametrize("B,Ns,use_mask", [(64, 64, False), (33, 192, True)])
def test_composite_bwd(cuda_lib, B, Ns, use_mask):
    from samplenerfro_b200 import ops
    raw, t, dirs, bk = _composite_inputs(B, Ns, 100 + Ns)
    gen = torch.Generator().manual_seed(2)
    mask = (torch.rand(B, Ns, generator=gen) > 0.3).float() if use_mask else None
    g_rgb, g_tr, g_trb = torch.randn(B, 3, generator=gen), torch.randn(B, 1, generator=gen), torch.randn(B, 3, generator=gen)
    raw64 = raw.double().requires_grad_(True); bk64 = bk.double().requires_grad_(True)
    ref = O.volumetric_rendering(O.rgb_act(raw64[..., :3]), O.sigma_act(raw64[..., 3:4]), t.double(), dirs.double(),
                                 False, O.rgb_act(bk64), mask.double() if use_mask else None)
    loss = (ref[0] * g_rgb.double()).sum() + (ref[5] * g_tr.double()).sum() + (ref[6] * g_trb.double()).sum()
    loss.backward()
    d_raw, d_bk = ops.composite_bwd(raw.cuda(), t.cuda(), dirs.cuda(), bk.cuda(), mask.cuda() if use_mask else None,
                                    g_rgb.cuda(), g_tr.reshape(-1).contiguous().cuda(), g_trb.cuda())
    assert H.rel_err(d_raw, raw64.grad) < 2e-5, H.rel_err(d_raw, raw64.grad)
    assert H.rel_err(d_bk, bk64.grad) < 2e-5, H.rel_err(d_bk, bk64.grad)


def _resample_setup(scene, B, randomized, weights_fn, S=768, P=12):
    from samplenerfro_b200 import ops
    Nc, Nf = 64, 128
    o, d = H.random_rays(B, seed=11)
    path = ops.march(scene["table_cu"], scene["ndim"], scene["nmin"], scene["nmax"], o.cuda(), d.cuda(), 2.0, 6.0, S)
    gen = torch.Generator().manual_seed(4)
    jit = torch.arange(0, S, P) + torch.randint(0, P, (Nc,), generator=gen)
    w = weights_fn(torch.rand(B, Nc, generator=gen))
    rp, rd, rt, _, rg = [x.cpu().contiguous() for x in ops.path_views(path)]
    t_c = rt[:, jit].contiguous()
    if randomized:
        u = O.stratified_u(torch.rand(B, Nf, generator=gen) * (1 / Nf - float(np.finfo(np.float32).eps)))
    else:
        u = O.deterministic_u(Nf)
    return path, (rp, rd, rt, rg), jit, t_c, w, u, Nf


@pytest.mark.parametrize("B,randomized", [(200, False), (129, True)])
def test_resample(cuda_lib, scene, B, randomized):
    """Well-conditioned pdf (every bin has mass): element-wise agreement with the oracle."""
    from samplenerfro_b200 import ops
    path, (rp, rd, rt, rg), jit, t_c, w, u, Nf = _resample_setup(scene, B, randomized, lambda r: 0.05 + r)
    w[0] = 0.0   # all-zero weights -> the 1e-5 padding branch gives a uniform pdf
    t_mid = 0.5 * (t_c[..., 1:] + t_c[..., :-1])
    z, pos, dirs, grads = O.sample_pdf(t_mid, w[..., 1:-1], rp, rd, rt, rg, u, jit)
    t_f, pos_f, dir_f, grad_f = ops.resample(path, t_c.cuda(), w.cuda(), u.cuda(), Nf, want_grad=True)
    assert (t_f.cpu() - z).abs().max().item() < 2e-5, (t_f.cpu() - z).abs().max().item()
    # the march-step choice is discontinuous in z: compare rows where both picked the same step (all but ulp-ties)
    same = (dir_f.cpu() == dirs).all(dim=-1)
    assert same.float().mean().item() > 0.995, same.float().mean().item()
    assert (pos_f.cpu() - pos)[same].abs().max().item() < 3e-5
    assert torch.equal(grad_f.cpu()[same], grads[same])
    assert (pos_f.cpu() - pos).abs().max().item() < 2e-4   # even at a tie the extrapolated point is continuous
    # the three search variants (t column staged in smem / strided two-level / compact records) agree bit for bit
    strided = ops.resample(ops.BentPath(path.rec, None), t_c.cuda(), w.cuda(), u.cuda(), Nf, want_grad=True)
    compact = ops.resample(ops.BentPath(path.rec[..., :8].contiguous(), path.t), t_c.cuda(), w.cuda(), u.cuda(), Nf)
    for a, b in zip((t_f, pos_f, dir_f, grad_f), strided):
        assert torch.equal(a, b)
    for a, b in zip((t_f, pos_f, dir_f), compact[:3]):
        assert torch.equal(a, b)


def test_resample_long_path(cuda_lib, scene):
    """S = 1536 (ball/glass/pen configs): the staged t column needs 24 KB of dynamic shared memory per CTA."""
    from samplenerfro_b200 import ops
    path, (rp, rd, rt, rg), jit, t_c, w, u, Nf = _resample_setup(scene, 96, False, lambda r: 0.05 + r, S=1536, P=24)
    t_mid = 0.5 * (t_c[..., 1:] + t_c[..., :-1])
    z, pos, dirs, grads = O.sample_pdf(t_mid, w[..., 1:-1], rp, rd, rt, rg, u, jit)
    t_f, pos_f, dir_f, _ = ops.resample(path, t_c.cuda(), w.cuda(), u.cuda(), Nf)
    assert (t_f.cpu() - z).abs().max().item() < 2e-5
    same = (dir_f.cpu() == dirs).all(dim=-1)
    assert same.float().mean().item() > 0.995
    assert (pos_f.cpu() - pos)[same].abs().max().item() < 3e-5


def test_resample_degenerate_pdf(cuda_lib, scene):
    """Nearly-empty bins make the inverse CDF ill-conditioned (a bin of mass 1e-8 maps an fp32 ulp of the cdf to
    the whole bin), so here the check is on invariants instead of values: sorted output, every coarse t present,
    and F(z) == u for the piecewise-linear CDF F evaluated in fp64."""
    from samplenerfro_b200 import ops
    B = 64
    path, (rp, rd, rt, rg), jit, t_c, w, u, Nf = _resample_setup(scene, B, False, lambda r: r ** 6)
    w[1, 10:] = 0.0       # cdf plateau at ~1
    w[2, :] = 0.0; w[2, 30] = 1.0   # a single occupied bin
    t_f, pos_f, dir_f, _ = ops.resample(path, t_c.cuda(), w.cuda(), u.cuda(), Nf)
    t_f = t_f.cpu()
    assert (t_f[:, 1:] >= t_f[:, :-1]).all()
    assert torch.isfinite(pos_f).all() and torch.isfinite(t_f).all()
    bins = (0.5 * (t_c[:, 1:] + t_c[:, :-1])).double()
    for r in range(B):
        rem = t_f[r].tolist()
        for v in t_c[r].tolist():
            rem.remove(v)           # raises if a coarse sample is missing from the merged list
        z = np.array(rem)
        ww = w[r, 1:-1].double().numpy()
        pad = max(0.0, 1e-5 - ww.sum()); ww = ww + pad / ww.size
        cdf = np.concatenate([[0.0], np.minimum(1.0, np.cumsum(ww / ww.sum()))]); cdf[-1] = 1.0
        assert z.min() >= bins[r, 0].item() - 1e-6 and z.max() <= bins[r, -1].item() + 1e-6
        Fz = np.interp(z, bins[r].numpy(), cdf)
        assert np.abs(Fz - u.double().numpy()).max() < 5e-6, (r, np.abs(Fz - u.double().numpy()).max())


def test_bkgd_mlp(cuda_lib):
    from samplenerfro_b200 import ops
    gen = torch.Generator().manual_seed(0)
    p = O.init_small_mlp(gen, bias_scale=0.1)
    B, Nc = 1000, 4
    d = torch.randn(B, Nc, 3, generator=gen)
    d = d / d.norm(dim=-1, keepdim=True)
    ref = O.small_mlp(p, O.pos_enc(d[:, -1:], 0, 4))[:, 0]
    w = ops.bkgd_pack(H.to_cuda_params(p))
    out = ops.bkgd_mlp_fwd(w, d.cuda().contiguous(), B, stride_floats=Nc * 3, offset_floats=(Nc - 1) * 3).cpu()
    assert (out - ref).abs().max().item() < 2e-5, (out - ref).abs().max().item()


@pytest.mark.parametrize("B", [1, 63, 64, 1000, 148 * 64 * 2 + 17])
def test_bkgd_mlp_on_the_tensor_pipe(cuda_lib, B):
    """The background MLP through the so3 evaluator (fp16 hi/lo split operands on tcgen05, fp32 accumulation): within 2e-6 of
    the fp32 CUDA-core kernel relative to the output scale, and within the same 2e-5 of the oracle; strided directions, ragged
    tiles, more tiles than SMs."""
    from samplenerfro_b200 import ops
    gen = torch.Generator().manual_seed(B)
    p = O.init_small_mlp(gen, bias_scale=0.1)
    Nc = 3
    d = torch.randn(B, Nc, 3, generator=gen)
    d = d / d.norm(dim=-1, keepdim=True)
    w = ops.bkgd_pack(H.to_cuda_params(p))
    dc = d.cuda().contiguous()
    fp32 = ops.bkgd_mlp_fwd(w, dc, B, stride_floats=Nc * 3, offset_floats=(Nc - 1) * 3)
    tc = ops.bkgd_mlp_fwd_tc(ops.bkgd_tc_pack(w), dc, B, stride_floats=Nc * 3, offset_floats=(Nc - 1) * 3)
    scale = fp32.abs().max().item()
    assert (tc - fp32).abs().max().item() < 2e-6 * max(scale, 1.0), ((tc - fp32).abs().max().item(), scale)
    ref = O.small_mlp(p, O.pos_enc(d[:, -1:], 0, 4))[:, 0]
    assert (tc.cpu() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("M", [128, 1000, 148 * 128 + 77, 60000, 74 * 512 * 3 + 333])
def test_encmlp_vs_bf16_oracle(cuda_lib, M):
    """Per-layer outputs vs an oracle that rounds operands to bf16 at the same points (fp32 accumulate).
    Stated bf16 tolerance: 2 bf16 ulps (2^-7 relative to the layer's max magnitude) per layer output."""
    from samplenerfro_b200 import ops
    gen = torch.Generator().manual_seed(M)
    p = O.init_nerf_mlp(gen, bias_scale=0.1)
    pos = (torch.rand(M, 3, generator=gen) * 2 - 1) * 3.0
    dirs = torch.randn(M, 3, generator=gen)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    packed = ops.encmlp_pack(H.to_cuda_params(p))
    raw, layers = ops.encmlp_fwd(packed, pos.cuda(), dirs.cuda(), debug_layers=True)
    torch.cuda.synchronize()
    raw2 = ops.encmlp_fwd(packed, pos.cuda(), dirs.cuda())
    # the production launch may be the CTA-pair kernel, which accumulates the k-blocks in a different order
    # (bf16 rounding of an activation can flip -> same 2-ulp tolerance as against the oracle)
    assert (raw - raw2).abs().max().item() <= 2.0 ** -7 * raw.abs().max().item(), "debug and production launches disagree"
    n_chk = min(M, 4096)
    sel = torch.randperm(M, generator=gen)[:n_chk]
    rgb_ref, sig_ref, lay_ref = O.nerf_mlp(p, O.pos_enc(pos[sel][:, None], 0, 10), O.pos_enc(dirs[sel][:, None], 0, 4),
                                           emulate_bf16=True, return_layers=True)
    layers = layers.float().cpu()
    for li, ref in enumerate(lay_ref):
        got = layers[li, sel, :ref.shape[1]]
        tol = 2.0 ** -7 * ref.abs().max().item()
        assert (got - ref).abs().max().item() <= tol, (li, (got - ref).abs().max().item(), tol)
    ref = torch.cat([rgb_ref[:, 0], sig_ref[:, 0]], dim=-1)
    err = (raw.cpu()[sel] - ref).abs().max().item()
    assert err < 2.0 ** -7 * ref.abs().max().item(), err
    err2 = (raw2.cpu()[sel] - ref).abs().max().item()
    assert err2 < 2.0 ** -7 * ref.abs().max().item(), err2
    # and against the true fp32 MLP: bf16 compute stays within ~1 % of the output scale
    rgb32, sig32 = O.nerf_mlp(p, O.pos_enc(pos[sel][:, None], 0, 10), O.pos_enc(dirs[sel][:, None], 0, 4))
    ref32 = torch.cat([rgb32[:, 0], sig32[:, 0]], dim=-1)
    assert (raw.cpu()[sel] - ref32).abs().max().item() < 0.03 * ref32.abs().max().item()
