"""Flax msgpack checkpoint interop: wire format, tree names, round trip, resume."""
import os

import msgpack
import numpy as np
import torch


def _variables():
    from samplenerfro_b200 import models
    gen = torch.Generator().manual_seed(0)
    return {"params": {"coarse_mlp": models.init_nerf_mlp_params(gen, "cpu"), "fine_mlp": models.init_nerf_mlp_params(gen, "cpu"),
                       "bkgd_mlp": models.init_small_mlp_params(gen, "cpu"),
                       "path_sampler": {"scan": {"idx_model": {"so3_mlp": models.init_small_mlp_params(gen, "cpu", in_dim=60, out_std=1e-5)}}}}}


def test_wire_format_is_flax_msgpack():
    """Bytes written here decode with plain msgpack into the structure flax.serialization documents: ndarray =
    ExtType(1, packb((shape, dtype.name, raw bytes)))."""
    from samplenerfro_b200 import checkpoint
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    blob = checkpoint.to_bytes({"step": np.asarray(7, dtype=np.int64), "params": {"params": {"m": {"Dense_0": {"kernel": a}}}}})
    raw = msgpack.unpackb(blob, raw=False)                       # no ext hook: ExtType objects stay visible
    ext = raw["params"]["params"]["m"]["Dense_0"]["kernel"]
    assert isinstance(ext, msgpack.ExtType) and ext.code == 1
    shape, dtype, buf = msgpack.unpackb(ext.data, raw=False)
    assert shape == [2, 3] and dtype == "float32" and np.array_equal(np.frombuffer(buf, np.float32).reshape(2, 3), a)
    # and a blob assembled the way flax does it (tuple shape, bin payload) restores here
    flax_like = msgpack.packb({"step": 3, "params": {"params": {"m": {"bias": msgpack.ExtType(
        1, msgpack.packb(((3,), "float32", np.float32([1, 2, 3]).tobytes()), use_bin_type=True))}}}}, use_bin_type=True)
    back = checkpoint.from_bytes(flax_like)
    assert back["step"] == 3 and np.array_equal(back["params"]["params"]["m"]["bias"], np.float32([1, 2, 3]))


def test_save_restore_round_trip_and_resume(tmp_path):
    from samplenerfro_b200 import checkpoint, train, utils
    V = _variables()
    state = train.TrainState.create(V, utils.Flags())
    state.step = 1234
    state.opt.count = 1234
    state.opt.mu.normal_(generator=torch.Generator().manual_seed(1))
    state.opt.nu.uniform_(generator=torch.Generator().manual_seed(2))
    ref = {k: v.clone() for k, v in (("theta", state.arena.theta), ("mu", state.opt.mu), ("nu", state.opt.nu))}
    p = checkpoint.save_checkpoint(str(tmp_path), state, 1234)
    assert os.path.basename(p) == "checkpoint_1234"
    checkpoint.save_checkpoint(str(tmp_path), state, 900)
    assert checkpoint.latest_checkpoint(str(tmp_path)).endswith("checkpoint_1234")      # natural order, like flax
    raw = checkpoint.restore_checkpoint(str(tmp_path), None)
    assert int(raw["step"]) == 1234
    k = raw["params"]["params"]["fine_mlp"]["Dense_5"]["kernel"]                         # the reference's access path
    assert k.shape == (319, 256) and np.array_equal(k, V["params"]["fine_mlp"]["Dense_5"]["kernel"].detach().numpy())
    assert raw["params"]["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]["Dense_3"]["kernel"].shape == (188, 128)
    # resume into a fresh state: parameters land IN the arena views, Adam moments and step come back
    V2 = _variables()
    for leaf in train.tree_leaves(V2):
        leaf.detach().zero_()
    s2 = train.TrainState.create(V2, utils.Flags())
    s2 = checkpoint.restore_checkpoint(str(tmp_path), s2)
    assert s2.step == 1234 and s2.opt.count == 1234
    assert torch.equal(s2.arena.theta, ref["theta"]) and torch.equal(s2.opt.mu, ref["mu"]) and torch.equal(s2.opt.nu, ref["nu"])
    assert V2["params"]["bkgd_mlp"]["Dense_0"]["kernel"].data_ptr() == s2.arena.theta_flat["bkgd_mlp"].data_ptr()
    # nothing to restore -> target returned unchanged (train.py:322 on a fresh run)
    assert checkpoint.restore_checkpoint(str(tmp_path / "empty"), s2) is s2


def test_keep_limit(tmp_path):
    from samplenerfro_b200 import checkpoint
    for step in range(5):
        checkpoint.save_checkpoint(str(tmp_path), {"step": step, "params": {"params": {}}}, step, keep=2)
    assert sorted(os.listdir(tmp_path)) == ["checkpoint_3", "checkpoint_4"]
