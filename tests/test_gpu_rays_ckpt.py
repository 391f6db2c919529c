"""Device ray generation / image assembly and checkpoint interop through the render path (rank 4)."""
import math

import numpy as np
import pytest
import torch

from oracle import rnerf_oracle as O
import rnerf_test_helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["blender", "opencv"])
@pytest.mark.parametrize("pc", [True, False])
def test_generate_rays_matches_reference_formula(cuda_lib, model, pc):
    """rnerf/datasets.py:216-242 / :486-518 as restated (numpy fp32) in the oracle: origins, directions and viewdirs
    bit-exact; radii within 2 ulp (numpy promotes the final division by np.sqrt(12) to float64 under NEP 50)."""
    from samplenerfro_b200 import synthetic, utils
    c2w = synthetic.camera_pose(0.9, 1.1, 4.03)
    h, w = 37, 53
    if model == "blender":
        focal = 0.5 * w / math.tan(0.5 * 0.6911112)
        ref = O.generate_rays(c2w, h, w, focal, use_pixel_centers=pc)
        got = utils.generate_rays(c2w, h, w, focal=focal, use_pixel_centers=pc)
    else:
        K = np.array([[61.5, 0, 25.3], [0, 60.25, 19.1], [0, 0, 1]], dtype=np.float32)
        ref = O.generate_rays(c2w, h, w, 0.0, use_pixel_centers=pc, opencv_K=K)
        got = utils.generate_rays(c2w, h, w, cam_mat=K, use_pixel_centers=pc)
    for nm in ("origins", "directions", "viewdirs"):
        assert torch.equal(getattr(got, nm).cpu(), getattr(ref, nm)), nm
    r, rr = got.radii.cpu(), ref.radii
    assert r.shape == rr.shape and ((r - rr).abs() <= 2.4e-7 * rr.abs()).all(), ((r - rr).abs() / rr.abs()).max()
    # row bands (multi-GPU partition of a frame) tile the full image exactly
    from samplenerfro_b200 import ops
    kw = dict(focal=focal) if model == "blender" else dict(cam_mat=K)
    band = ops.generate_rays(c2w, h, w, use_pixel_centers=pc, row0=30, n_rows=7, **kw)
    assert torch.equal(band[2], got.viewdirs[30:37]) and torch.equal(band[3], got.radii[30:37])


def test_render_view_equals_render_image_on_host_rays_and_checkpoint_round_trip(cuda_lib, tmp_path):
    """render_view (device rays + device assembly) == render_image over the host-built rays; the weights survive a
    Flax-format checkpoint written from one model and loaded into another (bit-identical render)."""
    from samplenerfro_b200 import checkpoint, models, synthetic, train, utils
    n, ndim, nmin, nmax = H.sphere_grid(G=24, radius=0.7, ws=3, sigma=1.0)
    args = utils.Flags(config="example", num_path_samples=12, white_bkgd=False, use_online_sparsity=False)
    model, variables = models.construct_nerf(4, None, args, ndim, nmin, nmax, n)
    c2w = synthetic.camera_pose(0.7, 1.0, 4.03)
    hh = ww = 24
    focal = 0.5 * ww / math.tan(0.5 * 0.6911112)

    def fn(vs):
        return lambda k0, k1, r: model.apply(vs, k0, k1, utils.namedtuple_map(lambda x: x.cuda(), r), False)

    with torch.no_grad():
        a = utils.render_view(fn(variables), c2w, hh, ww, 0, focal=focal, chunk=200)
        host = synthetic.blender_rays(c2w, hh, ww)
        b = utils.render_image(fn(variables), host, 0, False, chunk=200)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert abs(utils.image_psnr(a[0], a[0] * 0 + 0.5).item() - (-10 * math.log10(((a[0] - 0.5) ** 2).mean().item()))) < 1e-3
    # checkpoint -> fresh model with different weights -> same picture
    state = train.TrainState.create(variables, args)
    state.step = 42
    checkpoint.save_checkpoint(str(tmp_path), state, 42)
    model2, variables2 = models.construct_nerf(99, None, args, ndim, nmin, nmax, n)
    pre = checkpoint.restore_checkpoint(str(tmp_path), None)
    assert int(pre["step"]) == 42
    checkpoint.load_params_into(variables2, pre, names=["bkgd_mlp", "coarse_mlp", "fine_mlp"])      # eval.py:128-131
    with torch.no_grad():
        c = utils.render_view(lambda k0, k1, r: model2.apply(variables2, k0, k1, r, False), c2w, hh, ww, 0, focal=focal, chunk=200)
    assert torch.equal(c[0], a[0])
