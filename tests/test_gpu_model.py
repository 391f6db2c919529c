"""End-to-end parity of model.apply / render_image (CUDA path through the C ABI) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import rnerf_oracle as O
import rnerf_test_helpers as H

pytestmark = pytest.mark.gpu


def _flags(**kw):
    from samplenerfro_b200 import utils
    # configs/example.yaml values: Nc=64, Nf=128, P=12, near/far 2/6, white_bkgd false
    base = dict(config="example", num_coarse_samples=64, num_fine_samples=128, num_path_samples=12, white_bkgd=False,
                use_viewdirs=True, use_pixel_centers=True, randomized=True, use_online_sparsity=False,
                use_fine_sparsity=False, bg_weight=0.025, bg_smooth_weight=1.0, bg_patch_size=128)
    base.update(kw)
    return utils.Flags(**base)


def _oracle_vars(variables):
    def cv(t):
        if isinstance(t, dict):
            return {k: cv(v) for k, v in t.items()}
        return t.detach().cpu().clone()
    return cv(variables)


@pytest.fixture(scope="module")
def example_scene(cuda_lib):
    n, ndim, nmin, nmax = H.sphere_grid(G=48, radius=0.7, ws=3, sigma=1.0)
    return n, ndim, nmin, nmax


def test_model_apply_matches_oracle(cuda_lib, example_scene):
    """Config A shape (example.gin/yaml), 32x32 view, random-init weights: bent positions bit-exact,
    RGB >= 50 dB PSNR (north_star tolerance) against the fp32 oracle."""
    from samplenerfro_b200 import models, utils
    n, ndim, nmin, nmax = example_scene
    args = _flags()
    model, variables = models.construct_nerf(7, None, args, ndim, nmin, nmax, n)
    # non-zero biases so that every term of every layer is exercised
    gen = torch.Generator().manual_seed(1)
    for name in ("coarse_mlp", "fine_mlp", "bkgd_mlp"):
        for d in variables["params"][name].values():
            d["bias"].copy_(((torch.rand(d["bias"].shape, generator=gen) * 2 - 1) * 0.1).cuda())
    rays = H.camera_rays(32, 32, seed=2)
    flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]), rays)
    jitter = model.draw_jitter(11).cpu().long()
    u = O.deterministic_u(128)
    ret, loss_sp, dbg = model.apply(variables, 11, 12, utils.namedtuple_map(lambda r: r.cuda(), flat), False,
                                    jitter=jitter.int(), u=u, debug=True)
    cfg = O.ModelCfg(ndim=ndim, nmin=nmin, nmax=nmax, cfg_name="example")
    table = O.build_table(n, ndim, nmin, nmax)
    oret, _, odbg = O.nerf_model_apply(_oracle_vars(variables), table, cfg, O.Rays(*flat), jitter, u, debug=True)
    # 1. bent sample positions (1e-4 relative asked; bit-exact delivered)
    assert torch.equal(dbg["ray_pos"].cpu(), odbg["ray_pos"])
    assert torch.equal(dbg["ray_dir"].cpu(), odbg["ray_dir"])
    assert torch.equal(dbg["ray_dist"].cpu(), odbg["ray_dist"])
    assert (odbg["ray_dir"][:, -1] - flat.viewdirs).abs().max() > 1e-2, "scene does not refract"
    # 2. images
    for lvl in (0, 1):
        rgb, dist, acc, trans, trb = [x.cpu() for x in ret[lvl]]
        orgb, odist, oacc, otrans, otrb = oret[lvl]
        assert H.psnr(rgb, orgb) >= 50.0, (lvl, H.psnr(rgb, orgb))
        assert H.psnr(trb, otrb) >= 50.0
        assert (acc - oacc).abs().max() < 5e-3 and (trans - otrans).abs().max() < 5e-3
        assert (dist - odist).abs().max() < 2e-2
    assert float(loss_sp) == 0.0


def test_render_image_chunks_and_padding(cuda_lib, example_scene):
    """render_image (rnerf/utils.py:331-389): chunking / edge padding must not change any pixel.  Chunk sizes that
    run the same MLP kernel give bit-identical images; across the single-CTA / CTA-pair kernels (different fp32
    accumulation order of the k-blocks -> occasional bf16 rounding flips) the images agree to > 60 dB."""
    from samplenerfro_b200 import models, utils
    n, ndim, nmin, nmax = example_scene
    model, variables = models.construct_nerf(3, None, _flags(), ndim, nmin, nmax, n)
    rays = utils.namedtuple_map(lambda r: r.cuda(), H.camera_rays(20, 15, seed=5))
    fn = lambda k0, k1, r: model.apply(variables, k0, k1, r, False)
    rgb_a, dist_a, acc_a = utils.render_image(fn, rays, 0, False, chunk=150)
    rgb_b, dist_b, acc_b = utils.render_image(fn, rays, 0, False, chunk=77, world_size=8)
    assert rgb_a.shape == (20, 15, 3) and dist_a.shape == (20, 15, 1) and acc_a.shape == (20, 15, 1)
    assert torch.equal(rgb_a, rgb_b) and torch.equal(dist_a, dist_b) and torch.equal(acc_a, acc_b)
    rgb_c, dist_c, acc_c = utils.render_image(fn, rays, 0, False, chunk=8192)    # large chunk -> CTA-pair kernel
    assert H.psnr(rgb_c, rgb_a) > 60.0 and (acc_c - acc_a).abs().max() < 2e-3


def test_render_image_debug_variant(cuda_lib, example_scene):
    """The reference's commented debug variant (rnerf/utils.py:371-389, consumer extract_mesh.py:178): render_image returns
    the 8-tuple (rgb, distance, acc, ray_pos, ray_dir, idx_grad, trans, ray_pos_c) with the bent FINE samples and the coarse
    sample positions -- the "per-ray sample positions out" of the render-level API -- checked against the oracle."""
    from samplenerfro_b200 import models, utils
    n, ndim, nmin, nmax = example_scene
    model, variables = models.construct_nerf(3, None, _flags(), ndim, nmin, nmax, n)
    Hh, Ww = 12, 10
    rays = H.camera_rays(Hh, Ww, seed=5)
    jitter = model.draw_jitter(utils._split_key(0)[0])
    fn = lambda k0, k1, r: model.apply(variables, k0, k1, r, False, debug=True)
    out = utils.render_image(fn, utils.namedtuple_map(lambda r: r.cuda(), rays), 0, False, chunk=50, debug=True)
    assert len(out) == 8
    rgb, dist, acc, ray_pos, ray_dir, idx_grad, trans, ray_pos_c = [x.cpu() for x in out]
    assert ray_pos.shape == (Hh, Ww, 192, 3) and ray_dir.shape == (Hh, Ww, 192, 3) and idx_grad.shape == (Hh, Ww, 192, 3)
    assert trans.shape == (Hh, Ww, 1) and ray_pos_c.shape == (Hh, Ww, 64, 3) and rgb.shape == (Hh, Ww, 3)
    plain = utils.render_image(lambda k0, k1, r: model.apply(variables, k0, k1, r, False),
                               utils.namedtuple_map(lambda r: r.cuda(), rays), 0, False, chunk=50)
    assert len(plain) == 3 and H.psnr(plain[0], rgb) > 60.0
    flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]), rays)
    cfg = O.ModelCfg(ndim=ndim, nmin=nmin, nmax=nmax, cfg_name="example")
    oret, _, odbg = O.nerf_model_apply(_oracle_vars(variables), O.build_table(n, ndim, nmin, nmax), cfg, O.Rays(*flat),
                                       jitter.cpu().long(), O.deterministic_u(128), debug=True)
    assert torch.equal(ray_pos_c.reshape(-1, 64, 3), odbg["ray_pos_c"])                    # coarse samples: bit-exact
    # the fine samples sit where the coarse pass put its weights (bf16 MLP here, fp32 in the oracle), so they are compared
    # through the oracle's sample_pdf fed with THIS path's coarse weights: same tolerance as the resampling tests
    _, _, dbg = model.apply(variables, *utils._split_key(0), utils.namedtuple_map(lambda r: r.cuda(), flat), False, debug=True)
    t_c, w_c = dbg["t_c"].cpu(), dbg["weights_c"].cpu()
    t_f, pos_f, dir_f, grad_f = O.sample_pdf(0.5 * (t_c[:, 1:] + t_c[:, :-1]), w_c[:, 1:-1], odbg["ray_pos"], odbg["ray_dir"],
                                             odbg["ray_dist"], odbg["idx_grad"], O.deterministic_u(128), jitter.cpu().long())
    assert (ray_pos.reshape(-1, 192, 3) - pos_f).abs().max() < 1e-4
    assert (ray_dir.reshape(-1, 192, 3) - dir_f).abs().max() < 1e-4
    assert (idx_grad.reshape(-1, 192, 3) - grad_f).abs().max() < 1e-4
    assert idx_grad.abs().max() > 0.1, "no fine sample near the refractive boundary"
    assert (trans.reshape(-1) - oret[1][3].reshape(-1)).abs().max() < 1e-2
    with pytest.raises(ValueError):
        utils.render_image(lambda k0, k1, r: model.apply(variables, k0, k1, r, False),
                           utils.namedtuple_map(lambda r: r.cuda(), rays), 0, False, chunk=50, debug=True)


def test_bd_cut_dist_passes(cuda_lib):
    """ball.gin shape (S=1536, near/far 0.2/12, bd_cut_dist) against the oracle's extra composites (a15)."""
    from samplenerfro_b200 import models, utils
    n, ndim, nmin, nmax = H.sphere_grid(G=32, extent=2.0, radius=1.0, center=(0.0, 1.036, 0.0), cfg_name="ball", ws=5, sigma=3.0)
    args = _flags(config="ball", near=0.2, far=12.0, num_path_samples=24, bg_weight=1.0)
    args.gin_bindings = {"NerfModel": {"use_mask_bbox": False, "bd_cut_dist": 6.0}}
    model, variables = models.construct_nerf(5, None, args, ndim, nmin, nmax, n)
    o, d = H.random_rays(96, seed=3, radius=3.0, target_extent=0.8)
    o[:, 1] += 1.0
    rays = utils.Rays(o.cuda(), d.cuda(), d.cuda(), torch.ones(96, 1).cuda())
    jitter = model.draw_jitter(1).cpu().long()
    u = O.deterministic_u(128)
    ret, _ = model.apply(variables, 1, 2, rays, False, jitter=jitter.int(), u=u)
    cfg = O.ModelCfg(ndim=ndim, nmin=nmin, nmax=nmax, near=0.2, far=12.0, num_path_samples=24, bd_cut_dist=6.0, cfg_name="ball")
    oret, _ = O.nerf_model_apply(_oracle_vars(variables), O.build_table(n, ndim, nmin, nmax), cfg,
                                 O.Rays(o, d, d, torch.ones(96, 1)), jitter, u)
    for got, ref in zip(ret[1], oret[1]):
        assert (got.cpu() - ref).abs().max() < 1e-2, (got.cpu() - ref).abs().max()
    assert H.psnr(ret[1][0], oret[1][0]) >= 50.0


def test_unsupported_configs_fail_loudly(cuda_lib, example_scene):
    from samplenerfro_b200 import models
    n, ndim, nmin, nmax = example_scene
    with pytest.raises(NotImplementedError):
        models.construct_nerf(0, None, _flags(net_width=128), ndim, nmin, nmax, n)
    with pytest.raises(NotImplementedError):
        models.construct_nerf(0, None, _flags(legacy_posenc_order=True), ndim, nmin, nmax, n)


def test_cpu_tensors_are_rejected(cuda_lib):
    from samplenerfro_b200 import ops, _lib
    with pytest.raises(_lib.RnerfError):
        ops.grid_table(torch.ones(8), [2, 2, 2], [0.0] * 3, [1.0] * 3)


def test_all_stage_model_matches_oracle(cuda_lib):
    """stage="all" end to end (march with the so3 rotation -> coarse -> resample -> fine) vs the oracle, with so3 weights
    large enough to bend the rays visibly, annealed_alpha = 0.6."""
    from samplenerfro_b200 import models, utils
    n, ndim, nmin, nmax = H.sphere_grid(G=24, radius=0.7, ws=3, sigma=1.0)
    args = _flags(stage="all")
    model, variables = models.construct_nerf(5, None, args, ndim, nmin, nmax, n)
    so3 = variables["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]
    gen = torch.Generator().manual_seed(3)
    so3["Dense_4"]["kernel"].copy_((torch.randn(128, 3, generator=gen) * 0.05).cuda())
    so3["Dense_4"]["bias"].copy_((torch.randn(3, generator=gen) * 0.2).cuda())
    o, d = H.random_rays(64, seed=11)
    jitter = model.draw_jitter(2)
    u = O.deterministic_u(128)
    with torch.no_grad():
        ret, _, dbg = model.apply(variables, 0, 0, utils.Rays(o.cuda(), d.cuda(), d.cuda(), torch.ones(64, 1).cuda()), False, 0.6,
                                  jitter=jitter, u=u, debug=True)

    def cv(t):
        return {k: cv(v) for k, v in t.items()} if isinstance(t, dict) else t.detach().cpu()

    cfg = O.ModelCfg(ndim=ndim, nmin=nmin, nmax=nmax, cfg_name="example", stage="all")
    oret, _, odbg = O.nerf_model_apply(cv(variables), O.build_table(n, ndim, nmin, nmax), cfg, O.Rays(o, d, d, torch.ones(64, 1)),
                                       jitter.cpu().long(), u, annealed_alpha=0.6, debug=True)
    scale = odbg["ray_pos"].abs().max().item()
    assert (dbg["ray_pos"].cpu() - odbg["ray_pos"]).abs().max().item() < 1e-4 * scale
    plain_model, _ = models.construct_nerf(5, None, _flags(), ndim, nmin, nmax, n)
    with torch.no_grad():
        _, _, pdbg = plain_model.apply(variables, 0, 0, utils.Rays(o.cuda(), d.cuda(), d.cuda(), torch.ones(64, 1).cuda()), False,
                                       jitter=jitter, u=u, debug=True)
    assert (pdbg["ray_pos"] - dbg["ray_pos"]).abs().max().item() > 1e-3      # the rotation changed the paths
    assert H.psnr(ret[1][0], oret[1][0]) >= 50.0


def test_all_stage_march_many_active_rays_per_cta(cuda_lib):
    """A bundle of nearly parallel rays crosses the object boundary at the same steps, so all 128 rays of a CTA need the
    so3 MLP at once: exercises the multi-pass path of the compacted evaluation (64 columns per pass) and partial CTAs."""
    from samplenerfro_b200 import models, ops
    n, ndim, nmin, nmax = H.sphere_grid(G=24, radius=0.7, ws=3, sigma=1.0)
    gen = torch.Generator().manual_seed(5)
    B = 300
    o = torch.tensor([0.3, -3.9, 0.5]) + torch.randn(B, 3, generator=gen) * 0.01
    d = -o + torch.randn(B, 3, generator=gen) * 0.02
    d = d / d.norm(dim=-1, keepdim=True)
    model, variables = models.construct_nerf(5, None, _flags(stage="all"), ndim, nmin, nmax, n)
    so3 = variables["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]
    so3["Dense_4"]["kernel"].copy_((torch.randn(128, 3, generator=gen) * 0.05).cuda())
    so3["Dense_4"]["bias"].copy_((torch.randn(3, generator=gen) * 0.2).cuda())
    S = 768
    path = ops.march(model.table, ndim, nmin, nmax, o.cuda().contiguous(), d.cuda().contiguous(), 2.0, 6.0, S, bricks=model.bricks,
                     so3=(ops.so3_pack(so3), model.so3_window(0.8)))
    pos, dirs, dist, nn, g = ops.path_views(path)
    act = (g.norm(dim=-1) > 1e-3)
    assert act[:128].sum(dim=0).max().item() > 64          # more active rays in one CTA-step than one pass holds
    cpu = {k: {kk: vv.detach().cpu() for kk, vv in v.items()} for k, v in so3.items()}
    opos, odir, odist, _, _ = O.march(O.build_table(n, ndim, nmin, nmax), ndim, nmin, nmax, o, d, 2.0, 6.0, S, stage="all",
                                      so3_params=cpu, annealed_alpha=0.8)
    scale = opos.abs().max().item()
    assert (pos.cpu() - opos).abs().max().item() < 1e-4 * scale, (pos.cpu() - opos).abs().max().item()
    assert (dirs.cpu() - odir).abs().max().item() < 1e-4
    assert (dist.cpu() - odist).abs().max().item() < 1e-4 * odist.abs().max().item()


def test_all_stage_ragged_march_matches_lockstep(cuda_lib, monkeypatch):
    """Small launches of the "all"-stage march let every ray run at its own step and batch the so3 evaluations of rays at
    different steps (march_all_ragged_kernel).  Per ray the march arithmetic is unchanged; only the fp32 summation order
    inside so3_mlp depends on how many rays share an evaluation (row-split path for <= 8 columns), so records and the t
    column agree with the lockstep kernel to rounding (compact and full records, ragged ray count)."""
    from samplenerfro_b200 import models, ops
    n, ndim, nmin, nmax = H.sphere_grid(G=24, radius=0.7, ws=3, sigma=1.0)
    gen = torch.Generator().manual_seed(8)
    model, variables = models.construct_nerf(5, None, _flags(stage="all"), ndim, nmin, nmax, n)
    so3 = variables["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]
    so3["Dense_4"]["kernel"].copy_((torch.randn(128, 3, generator=gen) * 0.05).cuda())
    so3["Dense_4"]["bias"].copy_((torch.randn(3, generator=gen) * 0.2).cuda())
    o, d = H.random_rays(333, seed=4, target_extent=0.8)
    w = (ops.so3_pack(so3), model.so3_window(0.8))
    for compact in (True, False):
        monkeypatch.setenv("RNERF_SO3_RPC", "128")                       # lockstep kernel
        a = ops.march(model.table, ndim, nmin, nmax, o.cuda(), d.cuda(), 2.0, 6.0, 768, bricks=model.bricks, compact=compact, so3=w)
        monkeypatch.delenv("RNERF_SO3_RPC")                               # 333 rays -> 32 rays per CTA, ragged kernel
        b = ops.march(model.table, ndim, nmin, nmax, o.cuda(), d.cuda(), 2.0, 6.0, 768, bricks=model.bricks, compact=compact, so3=w)
        assert (a.rec - b.rec).abs().max().item() < 1e-5 * a.rec.abs().max().item() and (a.t - b.t).abs().max().item() < 1e-4
        homog = ops.path_views(a)[3][..., 0] == ops.path_views(a)[3][:, :1, 0]        # n still the ambient value
        first_bent = (~homog).float().argmax(dim=1)                                   # before the boundary: bit-identical
        k = torch.arange(768, device="cuda")[None]
        before = (k < first_bent[:, None]) | homog.all(dim=1, keepdim=True)
        assert torch.equal(a.rec[before], b.rec[before])
    act = (ops.path_views(b)[4].norm(dim=-1) > 1e-3)
    assert act.any(dim=1).sum().item() > 50                               # the MLP was needed on many rays


def test_ior_stage_is_the_reference_s_degenerate_stage(cuda_lib, example_scene):
    """stage="ior" (train.py:131-143): no rendering, loss_nrm = normal_loss = the constant 0.0 of
    rnerf/eikonal_utils.py:98 -- only weight decay reaches the parameters, and only path_sampler is trainable
    (train.py:295-301).  Rendering in this stage marches without the so3 rotation (rnerf/eikonal_utils.py:34 tests "all")."""
    from samplenerfro_b200 import models, train, utils
    n, ndim, nmin, nmax = example_scene
    args = _flags(stage="ior")
    args.weight_decay_mult, args.lr_delay_steps, args.normal_loss_weight = 0.1, 0, 1.0
    model, variables = models.construct_nerf(0, None, args, ndim, nmin, nmax, n)
    state = train.TrainState.create(variables, args)
    assert state.arena.buckets == ("path_sampler",)
    so3 = variables["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]["Dense_0"]["kernel"]
    coarse = variables["params"]["coarse_mlp"]["Dense_0"]["kernel"]
    assert so3.requires_grad and not coarse.requires_grad
    s0, c0 = so3.detach().clone(), coarse.detach().clone()
    batch = {"annealed_alpha": 0.5, **utils.GridPoints(model, 64).next_train()}
    state.step = 1
    state, stats, _ = train.train_step(model, 0, state, batch, args)
    assert float(stats["loss"]) == 0.0 and float(stats["loss_nrm"]) == 0.0
    assert torch.equal(coarse, c0)
    moved = so3.detach() - s0
    assert moved.abs().max().item() > 0 and (torch.sign(moved) == -torch.sign(s0))[s0 != 0].all()     # pure decay
    o, d = H.random_rays(16, seed=1)
    with torch.no_grad():
        ret, _ = model.apply(variables, 1, 2, utils.Rays(o.cuda(), d.cuda(), d.cuda(), torch.ones(16, 1).cuda()), False)
    assert torch.isfinite(ret[1][0]).all()


def test_so3_tensor_pipe_evaluator_matches_cuda_core_and_oracle(cuda_lib):
    """so3_mlp on tcgen05 kind::f16 with fp16 hi/lo split operands (csrc/so3_tc.cuh) against the fp32 CUDA-core chain and the oracle:
    VoxMLP.wrapper_grad_mlp (rnerf/ior_utils.py:225-267) on free-standing points.  Stated tolerance 2e-6 relative to the
    largest output (fp32-grade: 22-bit operands, the dropped lo*lo products are 2^-22 relative), ragged tile sizes included."""
    from samplenerfro_b200 import ops
    gen = torch.Generator().manual_seed(0)
    so3 = O.init_small_mlp(gen, in_dim=60, out_std=0.05)
    for d in so3.values():
        d["bias"] = (torch.rand(d["bias"].shape, generator=gen) * 2 - 1) * 0.05
    w = ops.so3_pack(H.to_cuda_params(so3))
    packed = ops.so3_tc_pack(w)
    assert packed.numel() == 16 * 16384          # 8 k-blocks of 64 x (hi, lo) x [128 neurons x 64 k] fp16
    for N, alpha in ((1, 1.0), (63, 0.35), (65, 0.72), (5000, 0.5)):
        pts = (torch.rand(N, 3, generator=gen) * 2 - 1) * 1.5
        cond = torch.randn(N, 3, generator=gen)
        window = [float(v) for v in O.cosine_easing_window(0, 9, 10, alpha * 10)]
        a = ops.so3_predict(w, window, pts.cuda(), cond.cuda()).cpu()
        b = ops.so3_predict_tc(packed, w, window, pts.cuda(), cond.cuda()).cpu()
        ref = O.so3_predict(so3, pts, cond, alpha)
        scale = ref.abs().max().item()
        assert torch.isfinite(b).all()
        assert (a - b).abs().max().item() < 2e-6 * scale, (N, (a - b).abs().max().item(), scale)
        assert (b - ref).abs().max().item() < 1e-5 * scale, (N, (b - ref).abs().max().item(), scale)
    # a device-resident window (what a captured training graph passes) gives the same numbers
    wd = torch.tensor(window, device="cuda", dtype=torch.float32)
    assert torch.equal(ops.so3_predict_tc(packed, w, wd, pts.cuda(), cond.cuda()).cpu(), b)


def test_all_stage_full_frame_march_on_the_tensor_pipe(cuda_lib, monkeypatch):
    """Launches large enough for lockstep CTAs (a frame) with compact records run so3_mlp on the tensor pipe
    (march_tc_kernel, 256 rays per CTA).  Against the CUDA-core kernel on every ray (same inputs, RNERF_SO3_TC=0) and
    against the oracle's scan on rays that cross the object: bent positions within the stated 1e-4 relative.  The batch
    starts with a bundle of nearly parallel rays (> 64 active rays in one CTA-step: the multi-pass path) and has a ragged
    tail (not a multiple of 256)."""
    from samplenerfro_b200 import models, ops
    n, ndim, nmin, nmax = H.sphere_grid(G=24, radius=0.7, ws=3, sigma=1.0)
    gen = torch.Generator().manual_seed(5)
    ob = torch.tensor([0.3, -3.9, 0.5]) + torch.randn(300, 3, generator=gen) * 0.01
    db = -ob + torch.randn(300, 3, generator=gen) * 0.02
    cam = H.camera_rays(112, 112, seed=4)
    o = torch.cat([ob, cam.origins.reshape(-1, 3)])[:12701].contiguous()
    d = torch.cat([db, cam.viewdirs.reshape(-1, 3)])[:12701]
    d = (d / d.norm(dim=-1, keepdim=True)).contiguous()
    model, variables = models.construct_nerf(5, None, _flags(stage="all"), ndim, nmin, nmax, n)
    so3 = variables["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]
    so3["Dense_4"]["kernel"].copy_((torch.randn(128, 3, generator=gen) * 0.05).cuda())
    so3["Dense_4"]["bias"].copy_((torch.randn(3, generator=gen) * 0.2).cuda())
    S = 768
    w = ops.so3_pack(so3)
    win = model.so3_window(0.8)
    args = (model.table, ndim, nmin, nmax, o.cuda(), d.cuda(), 2.0, 6.0, S)
    tc = ops.march(*args, bricks=model.bricks, compact=True, so3=(w, win), so3_tc=ops.so3_tc_pack(w))
    monkeypatch.setenv("RNERF_SO3_TC", "0")
    cc = ops.march(*args, bricks=model.bricks, compact=True, so3=(w, win), so3_tc=ops.so3_tc_pack(w))
    monkeypatch.delenv("RNERF_SO3_TC")
    full = ops.march(*args, bricks=model.bricks, compact=False, so3=(w, win))        # full records: CUDA-core chain, has grad n
    scale = cc.rec[..., 0:3].abs().max().item()
    assert torch.isfinite(tc.rec).all()
    err = (tc.rec[..., 0:3] - cc.rec[..., 0:3]).abs().max().item()
    assert err < 2e-5 * scale, err                                   # tensor pipe vs CUDA cores: summation order + 2^-21 products
    assert (tc.t - cc.t).abs().max().item() < 2e-5 * cc.t.abs().max().item()
    assert torch.equal(tc.rec[..., 3], tc.t)
    act = full.rec[..., 8:11].norm(dim=-1) > 1e-3
    assert act[:256].sum(dim=0).max().item() > 64                    # the first CTA needs more than one 64-column pass
    bent = act.any(dim=1).nonzero()[:, 0]
    assert bent.numel() > 2000
    pick = torch.cat([torch.arange(0, 300, 10), bent[torch.linspace(300, bent.numel() - 1, 34).long()].cpu()])
    cpu = {k: {kk: vv.detach().cpu() for kk, vv in v.items()} for k, v in so3.items()}
    opos, odir, odist, _, _ = O.march(O.build_table(n, ndim, nmin, nmax), ndim, nmin, nmax, o[pick], d[pick], 2.0, 6.0, S, stage="all",
                                      so3_params=cpu, annealed_alpha=0.8)
    assert (tc.rec[pick.cuda()][..., 0:3].cpu() - opos).abs().max().item() < 1e-4 * opos.abs().max().item()
    assert (ops.path_dirs(tc.rec[pick.cuda()].contiguous()).cpu() - odir).abs().max().item() < 1e-4
    assert (tc.t[pick.cuda()].cpu() - odist).abs().max().item() < 1e-4 * odist.abs().max().item()
    # the rotation is visible: the radiance-stage path differs
    plain = ops.march(*args, bricks=model.bricks, compact=True)
    assert (plain.rec[..., 0:3] - tc.rec[..., 0:3]).abs().max().item() > 1e-3


def test_all_stage_ragged_march_on_the_tensor_pipe(cuda_lib, monkeypatch):
    """Small launches (a training batch: rays out of lockstep, 32 rays per CTA) with the packed hi/lo image run so3_mlp on the
    tensor pipe too (march_ragged_tc_kernel).  Against the CUDA-core ragged kernel on the same inputs, compact and full
    records, and against the oracle's scan: same tolerances as the full-frame kernel."""
    from samplenerfro_b200 import models, ops
    n, ndim, nmin, nmax = H.sphere_grid(G=24, radius=0.7, ws=3, sigma=1.0)
    gen = torch.Generator().manual_seed(9)
    o, d = H.random_rays(777, seed=13, target_extent=0.6)
    model, variables = models.construct_nerf(5, None, _flags(stage="all"), ndim, nmin, nmax, n)
    so3 = variables["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]
    so3["Dense_4"]["kernel"].copy_((torch.randn(128, 3, generator=gen) * 0.05).cuda())
    so3["Dense_4"]["bias"].copy_((torch.randn(3, generator=gen) * 0.2).cuda())
    S = 768
    w = ops.so3_pack(so3)
    win = model.so3_window(0.8)
    args = (model.table, ndim, nmin, nmax, o.cuda(), d.cuda(), 2.0, 6.0, S)
    for compact in (True, False):
        tc = ops.march(*args, bricks=model.bricks, compact=compact, so3=(w, win), so3_tc=ops.so3_tc_pack(w))
        cc = ops.march(*args, bricks=model.bricks, compact=compact, so3=(w, win))
        scale = cc.rec[..., 0:3].abs().max().item()
        assert torch.isfinite(tc.rec).all()
        assert (tc.rec[..., 0:3] - cc.rec[..., 0:3]).abs().max().item() < 2e-5 * scale
        assert (tc.t - cc.t).abs().max().item() < 2e-5 * cc.t.abs().max().item()
        if not compact:
            assert torch.equal(tc.rec[..., 8:11], cc.rec[..., 8:11]) or (tc.rec[..., 8:11] - cc.rec[..., 8:11]).abs().max().item() < 1e-3
    cpu = {k: {kk: vv.detach().cpu() for kk, vv in v.items()} for k, v in so3.items()}
    opos, odir, odist, _, _ = O.march(O.build_table(n, ndim, nmin, nmax), ndim, nmin, nmax, o[:96], d[:96], 2.0, 6.0, S, stage="all",
                                      so3_params=cpu, annealed_alpha=0.8)
    assert (tc.rec[:96, :, 0:3].cpu() - opos).abs().max().item() < 1e-4 * opos.abs().max().item()
    assert (ops.path_dirs(tc.rec[:96].contiguous()).cpu() - odir).abs().max().item() < 1e-4
    plain = ops.march(*args, bricks=model.bricks, compact=True)
    assert (plain.rec[..., 0:3] - tc.rec[..., 0:3]).abs().max().item() > 1e-3
