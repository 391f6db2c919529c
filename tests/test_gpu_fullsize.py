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
