"""Analytic known-answer tests of the CPU oracle.  CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle import rnerf_oracle as O


def test_constant_grid_march_is_straight():
    G, n0, S = 6, 1.25, 40
    ndim, nmin, nmax = [G] * 3, [-1.0] * 3, [1.0] * 3
    table = O.build_table(torch.full((G ** 3, 1), n0), ndim, nmin, nmax)
    assert table[:, 1:].abs().max() == 0
    o = torch.tensor([[0.0, 0.0, 3.0], [1.0, -2.0, 0.5]])
    d = torch.tensor([[0.0, 0.0, -1.0], [-0.6, 0.8, 0.0]])
    pos, dirs, dist, n, g = O.march(table, ndim, nmin, nmax, o, d, 2.0, 6.0, S)
    step = 4.0 / (S - 1)
    k = torch.arange(S, dtype=torch.float64)[None, :, None]
    expect = o.double()[:, None] + 2.0 * d.double()[:, None] + k * (step / n0) * d.double()[:, None]
    assert (pos.double() - expect).abs().max() < 1e-5
    assert (dist.double() - (2.0 + k[..., 0] * step / n0)).abs().max() < 1e-5
    assert (dirs - d[:, None]).abs().max() < 1e-6
    assert (n - n0).abs().max() < 1e-6 and g.abs().max() == 0


def test_march_returns_state_before_each_step():
    """T4: ray_pos[:,0] = o + near d, and the emitted lookups are taken at the emitted positions."""
    G = 8
    ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
    gen = torch.Generator().manual_seed(0)
    table = O.build_table(1 + 0.3 * torch.rand(G ** 3, 1, generator=gen), ndim, nmin, nmax)
    o = torch.tensor([[0.1, 0.2, 3.0]]); d = torch.tensor([[0.0, 0.0, -1.0]])
    pos, dirs, dist, n, g = O.march(table, ndim, nmin, nmax, o, d, 2.0, 6.0, 24)
    assert torch.equal(pos[:, 0], o + 2.0 * d)
    look = O.linear3(table, ndim, nmin, nmax, pos[0])
    assert torch.equal(look[:, :1], n[0]) and torch.equal(look[:, 1:], g[0])
    assert (dirs.norm(dim=-1) - 1).abs().max() < 1e-6     # returned directions are normalised, the state is not


def test_trilinear_reproduces_corners_linear_fields_and_clamps():
    G = 5
    ndim, nmin, nmax = [G, G, G], [0.0, 0.0, 0.0], [4.0, 4.0, 4.0]
    lin = torch.arange(G, dtype=torch.float32)
    X, Y, Z = torch.meshgrid(lin, lin, lin, indexing="ij")
    f = (1.0 + 0.5 * X + 0.25 * Y - 0.125 * Z).reshape(-1, 1)
    table = O.build_table(f, ndim, nmin, nmax)
    corners = torch.stack([X, Y, Z], -1).reshape(-1, 3)
    assert torch.equal(O.linear3(table, ndim, nmin, nmax, corners), table)            # voxel corners: exact entries
    pts = torch.rand(200, 3, generator=torch.Generator().manual_seed(1)) * 4
    val = O.linear3(table, ndim, nmin, nmax, pts)[:, 0]
    assert (val - (1.0 + 0.5 * pts[:, 0] + 0.25 * pts[:, 1] - 0.125 * pts[:, 2])).abs().max() < 1e-5
    outside = torch.tensor([[-3.0, 2.0, 2.0], [9.0, 2.0, 2.0]])
    edge = torch.tensor([[0.0, 2.0, 2.0], [4.0, 2.0, 2.0]])
    assert torch.equal(O.linear3(table, ndim, nmin, nmax, outside), O.linear3(table, ndim, nmin, nmax, edge))


def test_gradient_table_interior_exact_boundary_half():
    """Central differences on an edge-padded grid: a linear ramp gives the exact slope inside, half at the faces."""
    G = 6
    ndim, nmin, nmax = [G] * 3, [0.0] * 3, [5.0] * 3
    lin = torch.arange(G, dtype=torch.float32)
    X, _, _ = torch.meshgrid(lin, lin, lin, indexing="ij")
    g = O.compute_grad((2.0 * X).reshape(-1, 1), ndim, nmin, nmax).reshape(G, G, G, 3)
    assert torch.all(g[1:-1, :, :, 0] == 2.0) and torch.all(g[0, :, :, 0] == 1.0) and torch.all(g[-1, :, :, 0] == 1.0)
    assert g[..., 1:].abs().max() == 0


def test_pos_enc_feature_order():
    x = torch.tensor([[0.1, 0.2, 0.3]])
    e = O.pos_enc(x, 0, 10)[0]
    assert e.shape == (63,)
    assert torch.equal(e[:3], x[0])
    for k in range(10):
        for c in range(3):
            assert abs(e[3 + 3 * k + c].item() - math.sin(2 ** k * x[0, c].item())) < 1e-5
            assert abs(e[33 + 3 * k + c].item() - math.cos(2 ** k * x[0, c].item())) < 2e-4
    assert O.pos_enc(x, 0, 4).shape[-1] == 27
    a = O.annealed_pos_enc(x[None], 0, 10, 10.0)[0, 0]      # alpha = max_deg -> window == 1
    assert a.shape == (60,)
    assert abs(a[0].item() - math.sin(0.1)) < 1e-6 and abs(a[3].item() - math.cos(0.1)) < 1e-6  # degree-major [sin xyz, cos xyz]


def test_compositing_invariants():
    B, Ns = 7, 33
    gen = torch.Generator().manual_seed(0)
    rgb = torch.rand(B, Ns, 3, generator=gen); t = 2 + torch.sort(torch.rand(B, Ns, generator=gen) * 4).values
    dirs = torch.randn(B, Ns, 3, generator=gen); bk = torch.rand(B, 3, generator=gen)
    sigma = torch.rand(B, Ns, 1, generator=gen) * 5
    c, dist, acc, w, alpha, trans, trb = O.volumetric_rendering(rgb, sigma, t, dirs, False, bk)
    assert (acc + trans[:, 0] - 1).abs().max() < 1e-5                      # telescoping: acc + T_end == 1
    assert (trb - trans * bk).abs().max() == 0
    c0, d0, a0, *_ = O.volumetric_rendering(rgb, torch.zeros_like(sigma), t, dirs, False, bk)
    assert torch.equal(c0, bk) and a0.abs().max() == 0                    # sigma == 0 -> rgb == bkgd
    assert torch.equal(d0, t[:, 0])                                       # NaN distance -> 0 -> clipped to t_0 (T12)
    cw, *_ = O.volumetric_rendering(rgb, sigma, t, dirs, True, None)
    cn, *_ = O.volumetric_rendering(rgb, sigma, t, dirs, False, None)
    assert (cw - (cn + 1 - acc[:, None])).abs().max() < 1e-6


def test_pdf_uniform_weights_give_uniform_samples_and_tie_rule():
    Nc, Nf, S = 16, 32, 64
    t = torch.linspace(2, 6, S)[None]
    jitter = torch.arange(0, S, S // Nc)
    t_c = t[:, jitter]
    bins = 0.5 * (t_c[:, 1:] + t_c[:, :-1])
    u = O.deterministic_u(Nf)
    z = O.sorted_piecewise_constant_pdf(bins, torch.ones(1, Nc - 2), u)
    expect = bins[0, 0] + u * (bins[0, -1] - bins[0, 0])
    assert (z[0] - expect).abs().max() < 1e-5
    # sample_pdf index rule (T14): idx = max(#{ray_dist < z} - 1, 0); a tie z == ray_dist[k] picks k-1
    pos = torch.stack([t[0], torch.zeros(S), torch.zeros(S)], -1)[None]
    dirs = torch.tensor([1.0, 0.0, 0.0]).expand(1, S, 3)
    zz, p, dd, gg = O.sample_pdf(bins, torch.ones(1, Nc - 2), pos, dirs, t, torch.zeros(1, S, 3), u, jitter)
    assert zz.shape == (1, Nc + Nf) and (zz[0, 1:] >= zz[0, :-1]).all()
    assert (p[0, :, 0] - zz[0]).abs().max() < 1e-5           # straight path: extrapolated point sits at distance z
    k = 8
    cnt = torch.searchsorted(t[0], t[0, k:k + 1], right=False)
    assert cnt.item() == k and max(cnt.item() - 1, 0) == k - 1


def test_learning_rate_decay_endpoints():
    lr0 = O.learning_rate_decay(0, 5e-4, 5e-6, 200000, 2500, 0.01)
    assert abs(lr0 - 5e-4 * 0.01 * 0) < 1e-12                # start_rate = clip(step, 0, 1) = 0 at step 0
    lr1 = O.learning_rate_decay(2500, 5e-4, 5e-6, 200000, 2500, 0.01)
    assert abs(lr1 - math.exp(math.log(5e-4) * (1 - 2500 / 200000) + math.log(5e-6) * 2500 / 200000)) < 1e-12
    assert abs(O.learning_rate_decay(200000, 5e-4, 5e-6, 200000, 2500, 0.01) - 5e-6) < 1e-12


def test_widened_sigmoid_and_shifted_softplus():
    x = torch.tensor([-100.0, 0.0, 100.0])
    assert torch.allclose(O.rgb_act(x), torch.tensor([-0.001, 0.5, 1.001]), atol=1e-6)
    assert abs(O.sigma_act(torch.tensor([1.0])).item() - math.log(2.0)) < 1e-6


def test_loss_gradient_flows_only_to_radiance_params():
    """T7: in the radiance stage the march has no trainable inputs; so3_mlp receives no gradient."""
    G = 8
    ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
    gen = torch.Generator().manual_seed(0)
    table = O.build_table(1 + 0.2 * torch.rand(G ** 3, 1, generator=gen), ndim, nmin, nmax)
    cfg = O.ModelCfg(ndim=ndim, nmin=nmin, nmax=nmax, num_coarse_samples=8, num_fine_samples=8, num_path_samples=2)
    V = O.init_variables(0)
    for p in O.tree_leaves(V):
        p.requires_grad_(True)
    o = torch.tensor([[0.0, 0.0, 4.0]] * 4); d = torch.tensor([[0.0, 0.05 * i, -1.0] for i in range(4)])
    d = d / d.norm(dim=-1, keepdim=True)
    env = torch.randn(4, 4, 3, generator=gen); env = env / env.norm(dim=-1, keepdim=True)
    loss, stats = O.train_loss(V, table, cfg, O.Rays(o, d, d, torch.ones(4, 1)), torch.rand(4, 3, generator=gen), env,
                               O.default_jitter(cfg), O.deterministic_u(8), 1.0)
    loss.backward()
    P = V["params"]
    assert P["fine_mlp"]["Dense_0"]["kernel"].grad.abs().sum() > 0
    assert P["coarse_mlp"]["Dense_0"]["kernel"].grad.abs().sum() > 0
    assert P["bkgd_mlp"]["Dense_0"]["kernel"].grad.abs().sum() > 0
    g = P["path_sampler"]["scan"]["idx_model"]["so3_mlp"]["Dense_0"]["kernel"].grad
    assert g is None or g.abs().max() == 0      # only the (zero-weighted) weight_l2 term touches it


def test_rodrigues_is_a_rotation_and_identity_at_zero_angle():
    """rnerf/ior_utils.py:300-306: pred = |g|_safe * R(raw) g/|g|_safe -- the norm is preserved, raw -> 0 leaves g unchanged
    (theta is clamped to 1e-3, e = raw/theta -> 0: cos(1e-3) g), and a quarter turn about z maps x to y."""
    gen = torch.Generator().manual_seed(0)
    g = torch.randn(50, 3, generator=gen) * 3
    raw = torch.randn(50, 3, generator=gen)
    pred = O.rodrigues_grad(raw, g)
    assert torch.allclose(pred.norm(dim=-1), g.norm(dim=-1), rtol=1e-5)
    same = O.rodrigues_grad(torch.zeros(50, 3), g)
    assert torch.allclose(same, g * math.cos(1e-3), rtol=1e-6, atol=1e-7)
    q = O.rodrigues_grad(torch.tensor([[0.0, 0.0, math.pi / 2]]), torch.tensor([[2.0, 0.0, 0.0]]))
    assert torch.allclose(q, torch.tensor([[0.0, 2.0, 0.0]]), atol=1e-6)


def test_all_stage_loss_reaches_so3_mlp_only_through_the_coarse_samples():
    """'all' stage: so3_mlp receives gradient from the training loss, and the fine samples carry none of it -- they are
    stop_gradient (rnerf/model_utils.py:406-411), like ray_dist (rnerf/eikonal_utils.py:120) -- which is what the CUDA
    reverse sweep relies on (it is fed d ray_pos_c / d ray_dir_c only)."""
    G = 12
    ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
    lin = torch.linspace(-1.5, 1.5, G)
    X, Y, Z = torch.meshgrid(lin, lin, lin, indexing="ij")
    n = (1.0 + 0.5 * torch.sigmoid((0.8 - (X ** 2 + Y ** 2 + Z ** 2).sqrt()) * 6)).reshape(-1, 1)
    table = O.build_table(n, ndim, nmin, nmax)
    cfg = O.ModelCfg(ndim=ndim, nmin=nmin, nmax=nmax, num_coarse_samples=8, num_fine_samples=8, num_path_samples=4, stage="all")
    gen = torch.Generator().manual_seed(1)
    V = O.init_variables(0)
    so3 = V["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]
    so3["Dense_4"]["kernel"] = torch.randn(128, 3, generator=gen) * 0.05
    for p in O.tree_leaves(V):
        p.requires_grad_(True)
    o = torch.tensor([[0.1 * i, 0.0, 4.0] for i in range(6)]); d = torch.tensor([[0.0, 0.03 * i, -1.0] for i in range(6)])
    d = d / d.norm(dim=-1, keepdim=True)
    rays, px = O.Rays(o, d, d, torch.ones(6, 1)), torch.rand(6, 3, generator=gen)
    jitter, u = O.default_jitter(cfg), O.deterministic_u(8)
    loss, _ = O.train_loss(V, table, cfg, rays, px, None, jitter, u, 0.7, bg_smooth_weight=0.0)
    loss.backward()
    g = V["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]["Dense_0"]["kernel"].grad
    assert g is not None and g.abs().max() > 0
    # the resampled (fine) positions / directions are cut from the graph although the path they come from is not
    ray_pos, ray_dir, ray_dist, _, idx_grad = O.march(table, ndim, nmin, nmax, o, d, cfg.near, cfg.far, 32, stage="all",
                                                      so3_params=V["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"],
                                                      annealed_alpha=0.7)
    jl = jitter.long()
    t_mid = 0.5 * (ray_dist[:, jl][..., 1:] + ray_dist[:, jl][..., :-1])
    w = torch.rand(6, 6, generator=gen)
    _, pos_f, dir_f, _ = O.sample_pdf(t_mid, w, ray_pos, ray_dir, ray_dist.detach(), idx_grad, u, jl)
    assert not pos_f.requires_grad and not dir_f.requires_grad
