"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous ray sharding and the bucketed gradient
all-reduce-mean that stands in for jax.lax.pmean(grads, "batch") (train.py:166)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from samplenerfro_b200 import train, utils
    gen = torch.Generator().manual_seed(0)
    full = torch.randn(8, 5, generator=gen)                       # the "global batch" every rank can reconstruct
    lo, hi = utils.shard_range(full.shape[0], rank, world)
    shard = full[lo:hi]
    variables = {"params": {name: {"Dense_0": {"kernel": torch.zeros(5, 3, requires_grad=True),
                                               "bias": torch.zeros(3, requires_grad=True)}}
                            for name in train.GRAD_BUCKETS}}
    variables["params"]["path_sampler"] = {"so3": {"kernel": torch.zeros(2, 2)}}
    w = torch.arange(15, dtype=torch.float32).reshape(5, 3)
    for name in train.GRAD_BUCKETS:                               # per-rank mean loss over its shard, like the reference
        d = variables["params"][name]["Dense_0"]
        loss = ((shard @ (d["kernel"] + w) + d["bias"]) ** 2).mean()
        loss.backward()
    train.allreduce_mean_grads(variables, world)
    # expected: mean over ranks of per-shard gradients == gradient of the global-batch mean (equal shard sizes)
    k = torch.zeros(5, 3, requires_grad=True); b = torch.zeros(3, requires_grad=True)
    ((full @ (k + w) + b) ** 2).mean().backward()
    ok = all(torch.allclose(variables["params"][n]["Dense_0"]["kernel"].grad, k.grad, atol=1e-5) and
             torch.allclose(variables["params"][n]["Dense_0"]["bias"].grad, b.grad, atol=1e-5) for n in train.GRAD_BUCKETS)
    q.put((rank, ok, (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_gradient_mean_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [(0, 4), (4, 8)]


def _arena_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from samplenerfro_b200 import models, train, utils
    gen = torch.Generator().manual_seed(0)
    V = {"params": {"coarse_mlp": models.init_nerf_mlp_params(gen, "cpu"), "fine_mlp": models.init_nerf_mlp_params(gen, "cpu"),
                    "bkgd_mlp": models.init_small_mlp_params(gen, "cpu"),
                    "path_sampler": {"so3_mlp": models.init_small_mlp_params(gen, "cpu", in_dim=60, out_std=1e-5)}}}
    state = train.TrainState.create(V, utils.Flags())
    arena = state.arena
    # each rank's "loss": (rank+1) * sum of every trainable leaf  ->  gradient (rank+1) everywhere, mean 1.5 at N=2
    loss = sum(p.sum() for n in train.GRAD_BUCKETS for p in train.tree_leaves(V["params"][n])) * (rank + 1)
    loss.backward()
    arena.allreduce_mean(world)
    lo, hi = arena.bucket_range["bkgd_mlp"]
    leaves = [p for n in train.GRAD_BUCKETS for p in train.tree_leaves(V["params"][n])]
    ok = all(bool(torch.allclose(p.grad, torch.full((1,), 1.5))) for p in leaves)
    ok = ok and bool(torch.allclose(arena.grad[lo:hi], torch.full((1,), 1.5)))          # the bkgd bucket is dense
    ok = ok and abs(arena.grad.sum().item() - 1.5 * sum(p.numel() for p in leaves)) < 1.0   # padding stayed zero
    ok = ok and V["params"]["fine_mlp"]["Dense_5"]["kernel"].grad.data_ptr() == arena.sinks["fine_mlp"][10].data_ptr()
    ok = ok and V["params"]["path_sampler"]["so3_mlp"]["Dense_0"]["kernel"].grad is None
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_arena_allreduce():
    """ParamArena: gradients accumulate into the flat buffer's views and the per-bucket in-place all-reduce-mean over
    gloo equals pmean (train.py:166)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_arena_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]


def _arena_all_stage_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from samplenerfro_b200 import models, train, utils
    gen = torch.Generator().manual_seed(0)
    so3 = models.init_small_mlp_params(gen, "cpu", in_dim=60, out_std=1e-5)
    want_image = torch.cat([so3[f"Dense_{i}"]["kernel"].reshape(-1) for i in range(5)] +
                           [so3[f"Dense_{i}"]["bias"].reshape(-1) for i in range(5)]).clone()
    V = {"params": {"coarse_mlp": models.init_nerf_mlp_params(gen, "cpu"), "fine_mlp": models.init_nerf_mlp_params(gen, "cpu"),
                    "bkgd_mlp": models.init_small_mlp_params(gen, "cpu"),
                    "path_sampler": {"scan": {"idx_model": {"so3_mlp": so3}}}}}
    state = train.TrainState.create(V, utils.Flags(stage="all"))
    arena = state.arena
    names = train.ALL_STAGE_BUCKETS
    ok = arena.buckets == names and len(arena.bucket_grads()) == 4
    # the so3 bucket is the march kernels' weight image (5 kernels, then 5 biases), dense, and its gradient view is the sink
    ok = ok and bool(torch.equal(arena.theta_flat["so3_mlp"], want_image))
    ok = ok and arena.sinks["so3_mlp"].numel() == want_image.numel() == 64896 + 4 * 128 + 3
    leaf = V["params"]["path_sampler"]["scan"]["idx_model"]["so3_mlp"]["Dense_3"]["kernel"]
    ok = ok and leaf.requires_grad and leaf.grad.data_ptr() == arena.sinks["so3_mlp"][7680 + 2 * 16384:].data_ptr()
    # a kernel that accumulates into the sink (the reverse sweep does) is seen by the leaves and by the all-reduce
    arena.sinks["so3_mlp"].add_(float(rank + 1))
    loss = sum(p.sum() for n in train.GRAD_BUCKETS for p in train.tree_leaves(V["params"][n])) * (rank + 1)
    loss.backward()
    arena.allreduce_mean(world)
    leaves = [p for n in names for p in train.tree_leaves(V["params"][n])]
    ok = ok and all(bool(torch.allclose(p.grad, torch.full((1,), 1.5))) for p in leaves)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_arena_allreduce_all_stage():
    """"all" stage (train.py:302-310): so3_mlp is a fourth bucket laid out as the march kernels' weight image; gradients a
    kernel accumulates into its sink are averaged over ranks like the others."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_arena_all_stage_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]


def _fake_render_fn(key_0, key_1, rays):
    """Stands in for model.apply on CPU: any per-ray function will do (rays are independent)."""
    rgb = torch.sin(rays.origins * 3.0 + rays.viewdirs)
    dist_ = (rays.origins * rays.viewdirs).sum(-1)
    acc = rays.viewdirs.abs().sum(-1)
    return [(rgb, dist_, acc, acc[:, None], rgb)], torch.zeros(())


def _sharded_render_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from samplenerfro_b200 import utils
    gen = torch.Generator().manual_seed(0)
    ok = True
    for (H, W) in ((5, 7), (1, 4), (8, 3)):                       # H not divisible by N; fewer rows than ranks; even split
        rays = utils.Rays(torch.randn(H, W, 3, generator=gen), torch.randn(H, W, 3, generator=gen),
                          torch.randn(H, W, 3, generator=gen), torch.rand(H, W, 1, generator=gen))
        want = utils.render_image(_fake_render_fn, rays, 0, True, chunk=6)
        got = utils.render_image_sharded(_fake_render_fn, rays, 0, True, chunk=6, rank=rank, world_size=world)
        ok = ok and all(bool(torch.equal(a, b)) for a, b in zip(want, got))
        r0, r1, per = utils.band_rows(H, rank, world)
        ok = ok and 0 <= r0 <= r1 <= H and per * world >= H
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_sharded_render_gathers_the_frame():
    """Render partitioning: row bands per rank, no data-path collective, bands all-gathered like eval.py:96;
    the assembled frame equals the single-process render_image bit for bit, ragged splits included."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_render_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
