"""BASELINE.json configs[1] at FULL size (800x800 frame = 640 000 rays, S=768, IoR grid 512^3, 64 + 192 samples) through
size-independent properties, plus the oracle on a sampled subset of rays (rays are independent, so the rows of the full
run must equal a run of the subset alone, and the subset is small enough for the CPU oracle)."""
import numpy as np
import pytest
import torch

from oracle import rnerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ship(cuda_lib):
    from samplenerfro_b200 import models, ops, synthetic, utils
    G = 512
    ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
    data = synthetic.ellipsoid_occupancy(G, 1.5, (1.0, 0.4, 0.6), ss=4)
    n = ops.grid_blur(synthetic.rescale_ior(data, "ship_skydome"), ndim, 9, 3.0)
    del data
    flags = utils.Flags(config="ship_skydome", num_path_samples=12, white_bkgd=False, use_online_sparsity=False)
    model, variables = models.construct_nerf(0, None, flags, ndim, nmin, nmax, n)
    del n
    rays = utils.generate_rays(synthetic.camera_pose(0.7, 1.0, 4.03), 800, 800, focal=0.5 * 800 / np.tan(0.5 * 0.6911112))
    flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]).contiguous(), rays)
    return model, variables, flat, (ndim, nmin, nmax)


def test_full_frame_properties_and_chunk_invariance(ship):
    from samplenerfro_b200 import utils
    model, variables, flat, _ = ship
    B = flat.origins.shape[0]
    assert B == 640000
    jitter = model.draw_jitter(3)
    with torch.no_grad():
        ret, _ = model.apply(variables, 1, 2, flat, False, jitter=jitter)
        rgb, dist, acc, trans, trb = ret[1]
        # compositing invariants on every ray of the frame
        assert (acc + trans[:, 0] - 1).abs().max().item() < 2e-5                    # sum w + T_end == 1 (telescoping)
        assert rgb.min().item() >= -0.0011 and rgb.max().item() <= 1.0011           # widened sigmoid range
        assert torch.isfinite(rgb).all() and torch.isfinite(dist).all()
        assert (dist >= 2.0 - 1e-4).all() and (dist <= 6.5).all()                    # clipped to [t_0, t_last] of its ray
        assert (trb >= -0.0011 * trans).all() and (trb <= 1.0011 * trans + 1e-7).all()       # T_end * bkgd colour
        # chunk invariance: the frame in five 128 000-ray pieces is bit-identical to the frame in one piece
        parts = []
        for i in range(0, B, 128000):
            r = utils.namedtuple_map(lambda x: x[i:i + 128000], flat)
            parts.append(model.apply(variables, 1, 2, r, False, jitter=jitter)[0][1][0])
        assert torch.equal(torch.cat(parts), rgb)
        # an arbitrary subset of rays rendered alone (ragged size) agrees to fp32 accumulation order: a row's position in
        # the batch decides which tile pair of the MLP kernel it joins, and the two pairs consume the encoding k-block of
        # layers 5 and 9 at different ends of the K loop
        idx = torch.randperm(B, generator=torch.Generator().manual_seed(0))[:1037].cuda()
        sub = utils.namedtuple_map(lambda x: x[idx].contiguous(), flat)
        rgb_sub = model.apply(variables, 1, 2, sub, False, jitter=jitter)[0][1][0]
        assert (rgb_sub - rgb[idx]).abs().max().item() < 1e-4


def test_full_size_march_sortedness_and_oracle_on_a_subset(ship):
    from samplenerfro_b200 import ops
    model, variables, flat, (ndim, nmin, nmax) = ship
    B, S = flat.origins.shape[0], 768
    path = ops.march(model.table, ndim, nmin, nmax, flat.origins, flat.viewdirs, 2.0, 6.0, S, bricks=model.bricks, compact=True)
    t = path.t
    assert (t[:, 1:] > t[:, :-1]).all()                                               # ray_dist strictly increasing
    assert (t[:, 0] == 2.0).all()
    step = (6.0 - 2.0) / (S - 1)
    # |p' - p| = step * |v| / n with n in [1, 1.5] and |v| tracking n between 1 and 1.5 (eikonal): the increments stay
    # within a factor 1.5 of the step size either way
    dt = t[:, 1:] - t[:, :-1]
    assert dt.max().item() <= step * 1.5 * 1.02 and dt.min().item() >= step / 1.5 * 0.98
    assert torch.equal(path.rec[..., 3], t)                                           # dense t column == record field
    n_rec = path.rec[..., 7]
    # the 729-tap fp32 blur of a constant region is the constant times the rounded kernel sum (1 - 2.7e-6 here)
    assert n_rec.min().item() >= 1.0 - 2e-5 and n_rec.max().item() <= 1.5 + 2e-5
    # oracle on 24 rays that cross the object, with the device-built 512^3 table (its construction is parity-tested at
    # small sizes): bit-exact rows
    bent = ((path.rec[:, -1, 4:7] - flat.viewdirs).norm(dim=-1) > 1e-2).nonzero()[:, 0]
    assert bent.numel() > 1000
    pick = bent[torch.linspace(0, bent.numel() - 1, 24).long()]
    table = model.table.cpu()
    opos, odir, odist, on, og = O.march(table, ndim, nmin, nmax, flat.origins[pick].cpu(), flat.viewdirs[pick].cpu(), 2.0, 6.0, S)
    assert torch.equal(path.rec[pick][..., 0:3].cpu(), opos)
    assert torch.equal(path.rec[pick][..., 3].cpu(), odist)
    assert torch.equal(ops.path_dirs(path.rec[pick].contiguous()).cpu(), odir)
    assert torch.equal(path.rec[pick][..., 7].cpu(), on[..., 0])


def test_config_d_ball_full_size_properties(cuda_lib):
    """BASELINE.json configs[3] at FULL size: ball.gin / ball.yaml shape -- 1008x756 OpenCV camera (762 048 rays), S = 1536
    (P = 24), near/far 0.2/12, G = 256 extent 2, blur 5/3, bd_cut_dist = 6 with the hard-coded ball box
    (rnerf/models.py:489-491) -- through size-independent properties, plus the oracle on a subset of rays that cross the ball.
    Also the row-band decomposition used to shard this frame over 8 GPUs: 8 bands rendered one after the other on this GPU
    (each from rays generated for its band only) reproduce the frame bit for bit."""
    from samplenerfro_b200 import models, ops, synthetic, utils
    G = 256
    ndim, nmin, nmax = [G] * 3, [-2.0] * 3, [2.0] * 3
    data = synthetic.ellipsoid_occupancy(G, 2.0, (1.0, 1.0, 1.0), center=(0.0, 1.036, 0.0), ss=4)
    n = ops.grid_blur(synthetic.rescale_ior(data, "ball"), ndim, 5, 3.0)
    flags = utils.Flags(config="ball", num_path_samples=24, white_bkgd=False, use_online_sparsity=False, near=0.2, far=12.0)
    flags.gin_bindings = {"NerfModel": {"use_mask_bbox": False, "bd_cut_dist": 6.0}}
    model, variables = models.construct_nerf(0, None, flags, ndim, nmin, nmax, n)
    Hh, Ww, S = 756, 1008, 1536
    assert model.num_march_steps == S
    c2w = np.eye(4); c2w[:3, 3] = [0.0, 1.036, -5.0]              # OpenCV camera 5 units in front of the ball (+z forward)
    K = np.array([[1100.0, 0, Ww / 2], [0, 1100.0, Hh / 2], [0, 0, 1]])
    rays = utils.generate_rays(c2w, Hh, Ww, cam_mat=K)
    flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]).contiguous(), rays)
    B = flat.origins.shape[0]
    assert B == 762048
    jitter = model.draw_jitter(3)
    fn = lambda k0, k1, r: model.apply(variables, k0, k1, r, False, jitter=jitter)
    with torch.no_grad():
        chunk = 190512                                                             # a quarter of the frame per call
        rgb, dist, acc = utils.render_image(fn, rays, 0, False, chunk=chunk)
        assert rgb.shape == (Hh, Ww, 3) and torch.isfinite(rgb).all() and torch.isfinite(dist).all()
        assert rgb.min().item() >= -0.0011 and rgb.max().item() <= 1.0011
        assert (acc >= -1e-6).all() and (acc <= 1 + 1e-5).all()
        # one quarter again with the full tuple: telescoping sum, bd_cut_dist outputs in range
        q = utils.namedtuple_map(lambda x: x[chunk:2 * chunk], flat)
        ret, _ = model.apply(variables, 1, 2, q, False, jitter=jitter)
        r1, d1, a1, tr1, trb1 = ret[1]
        r0, d0, a0, tr0, trb0 = ret[0]
        assert (a0 + tr0[:, 0] - 1).abs().max().item() < 2e-5                    # coarse level: plain composite
        assert (tr1 >= -1e-6).all() and (tr1 <= 1 + 1e-6).all()                   # masked transmittance (rnerf/models.py:504-513)
        assert (trb1 >= -0.0011 - 1e-6).all() and (trb1 <= 1.0011 + 1e-6).all()   # trans * (colour behind the box)
        assert (tr1[:, 0] >= 1 - a1 - 2e-5).all()       # masking samples out can only raise the transmittance above 1 - acc
        assert torch.equal(rgb.reshape(-1, 3)[chunk:2 * chunk], r1)
        # 8 row bands, each from its own device-generated rays = the frame (what render_view_sharded does on 8 GPUs)
        bands = []
        for rk in range(8):
            b0, b1, per = utils.band_rows(Hh, rk, 8)
            o, d, v, rr = ops.generate_rays(c2w, Hh, Ww, cam_mat=K, row0=b0, n_rows=b1 - b0)
            bands.append(utils.render_image(fn, utils.Rays(o, d, v, rr), 0, False, chunk=chunk)[0])
        assert torch.equal(torch.cat(bands, dim=0), rgb)
        # the march at S = 1536: sortedness, start value, oracle rows on rays that cross the ball
        sub = utils.namedtuple_map(lambda x: x[::37].contiguous(), flat)
        path = ops.march(model.table, ndim, nmin, nmax, sub.origins, sub.viewdirs, 0.2, 12.0, S, bricks=model.bricks, compact=True)
        t = path.t
        assert (t[:, 1:] > t[:, :-1]).all() and (t[:, 0] == np.float32(0.2)).all()
        bent = ((path.rec[:, -1, 4:7] - sub.viewdirs).norm(dim=-1) > 1e-2).nonzero()[:, 0]
        assert bent.numel() > 500
        pick = bent[torch.linspace(0, bent.numel() - 1, 12).long()]
        opos, odir, odist, on, og = O.march(model.table.cpu(), ndim, nmin, nmax, sub.origins[pick].cpu(), sub.viewdirs[pick].cpu(),
                                            0.2, 12.0, S)
        assert torch.equal(path.rec[pick][..., 0:3].cpu(), opos) and torch.equal(path.rec[pick][..., 3].cpu(), odist)
        assert torch.equal(ops.path_dirs(path.rec[pick].contiguous()).cpu(), odir)
