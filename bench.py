#!/usr/bin/env python
"""Headline benchmark: rays/s of the refractive rendering hot path (march + enc/MLP + composite).

Workload (BASELINE.json configs[1]): ship_skydome synthetic scene, full 800x800 refractive render, S=768
march steps through a 512^3 IoR grid, 64 coarse + 192 fine radiance-MLP evaluations per ray, random-init
weights, synthetic blob -- one "step" is one full frame (640 000 rays).

  python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                   # CPU baseline (oracle port of the JAX path)

Prints ONE JSON line (see DESIGN.md "Measurement").

N = 1: `value` = one frame with the rays resident in HBM; `e2e` = the same frame through `utils.render_image` from pinned
HOST rays (H2D + D2H inside the timed region).
N > 1 (torchrun, one rank per GPU): STRONG scaling of the same ONE 800x800 frame -- the frame's rows are cut into N bands
(`utils.render_image_sharded`), every rank renders its band and the bands are all-gathered over NCCL inside the timed
region (the role of `jax.lax.all_gather` in eval.py:96-97); `e2e` adds the H2D of every rank's band rays and the D2H of the
assembled frame on rank 0.  Rank 0 reports frame rays / max-over-ranks device time.
Every N also carries a `train_step` key: the 4096-ray GLOBAL batch of BASELINE.json configs[2] (forward + backward +
bucketed NCCL all-reduce of the gradients + Adam, CUDA-graph replayed), B/N rays per rank -- train.py:166-167.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOP_PER_SAMPLE = 2 * 593408          # un-padded MACs of NerfMLP
# DRAM traffic of the enc+MLP kernel from its `ncu --set full` capture (profiles/r1a_encmlp_pair_kernel_ncu_summary.txt:
# dram__bytes_read 102.95 MB + dram__bytes_write 37.99 MB for a 4 194 304-sample launch) = 33.6 B/sample, against
# 40 B/sample algorithmic (24 B pos+dir in, 16 B raw out; part of the input is still L2-resident from its producer)
MLP_DRAM_BYTES_PER_SAMPLE = (102.945536e6 + 37.994752e6) / 4194304
# dram__bytes_read.sum + dram__bytes_write.sum of the fine-pass launch from `ncu --set full` at the bench's own launch size
# (samples per launch -> bytes); filled from profiles/<tag>_render_ncu_summary.txt
MLP_DRAM_BYTES = {122880000: 4915326000,     # profiles/r4a_render_ncu_summary.txt: fine launch of the 640 000-ray frame (40.0 B/sample)
                  40960000: 1617748000}      # the coarse launch (39.5 B/sample); algorithmic: 24 B in + 16 B out per sample
NC, NF, P = 64, 128, 12
S = NC * P
NEAR, FAR = 2.0, 6.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--side", type=int, default=800, help="image side (800 -> 640 000 rays per step)")
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--chunk", type=int, default=640000, help="rays per model.apply call (default: the whole 800x800 frame; the compact bent path of a frame is 17.7 GB of the 180 GB HBM)")
    ap.add_argument("--cpu-rays", type=int, default=16384, help="rays of the bounded CPU-baseline sample (~15-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the train_step key (configs[2])")
    ap.add_argument("--train-batch", type=int, default=4096, help="GLOBAL rays per optimisation step")
    ap.add_argument("--train-steps", type=int, default=20)
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(n_rays: int, grid_side: int = 96):
    """Oracle port of the reference's JAX path on the host cores: `n_rays` rays of the same camera through the
    same pipeline (S=768, 64+192 samples).  The grid is a smaller blurred blob (the march cost per step does not
    depend on the grid side on CPU; building a 512^3 table in numpy would dominate the sample)."""
    from oracle import rnerf_oracle as O
    import numpy as np
    G = grid_side
    lin = np.linspace(-1.5, 1.5, G)
    X, Y, Z = np.meshgrid(lin, lin, lin, indexing="ij")
    data = np.where((X / 1.0) ** 2 + (Y / 0.4) ** 2 + (Z / 0.6) ** 2 < 1.0, 1.33, 1.0).reshape(-1, 1)
    ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
    n = O.conv3d_normal(O.ior_rescale(data, "ship"), ndim, 3, 1.0)
    table = O.build_table(n, ndim, nmin, nmax)
    from samplenerfro_b200 import synthetic
    side = max(8, int(math.sqrt(n_rays)))
    rays = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), side, side)
    flat = [r.reshape(-1, r.shape[-1])[:n_rays] for r in rays]
    n_rays = flat[0].shape[0]
    V = O.init_variables(0)
    cfg = O.ModelCfg(ndim=ndim, nmin=nmin, nmax=nmax, cfg_name="ship")
    jitter, u = O.default_jitter(cfg), O.deterministic_u(NF)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.nerf_model_apply(V, table, cfg, O.Rays(*flat), jitter, u)
    dt = time.perf_counter() - t0
    return n_rays / dt, n_rays, dt


def run_reference(a, rank, world):
    """`--impl reference`: the reference's own implementation is JAX (not installable here), so this arm
    times the oracle port of it on the host cores; each step is a bounded sample of the same workload."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    n = min(a.cpu_rays, 4096)          # per step; K steps + warm-up stay within a couple of minutes
    for _ in range(min(a.warmup, 1)):
        cpu_baseline(min(n, 128))
    vals, t_all = [], 0.0
    for _ in range(a.steps):
        v, nr, dt = cpu_baseline(n)
        vals.append(v); t_all += dt
    value = n * a.steps / t_all
    line = {"impl": "reference", "metric": "rays/sec (march+MLP+composite)", "value": value, "unit": "rays/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_all / a.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ship_skydome 800x800 refractive render, S=768, 64 coarse + 192 fine samples/ray",
                       "sample": f"{n} rays per step of the same camera/pipeline on CPU"},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{n} rays/step x {a.steps} steps, oracle port (torch-CPU fp32) of the JAX path"},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ CUDA arm
def train_arm(a, model, rank, world, local, dev, barrier, peaks):
    """BASELINE.json configs[2]: the 4096-ray GLOBAL training batch (ship scene, S=768, G=512, 64+192 samples,
    bg_weight 0.025, bg_smooth 1.0 on a 128x128 env patch), B/N rays per rank; forward + backward + bucketed gradient
    all-reduce (jax.lax.pmean, train.py:166-167) + Adam, replayed from a CUDA graph.  Strong scaling.  Also times the same
    per-rank shard WITHOUT the collectives (world_size=1 semantics) so the exposed communication can be read off."""
    from samplenerfro_b200 import _lib, train, utils, synthetic
    targs = utils.Flags(config="ship_skydome-bkgd_no-partial-reflect_cycles", num_path_samples=P, white_bkgd=False,
                        use_online_sparsity=False, bg_weight=0.025, bg_smooth_weight=1.0, bg_patch_size=128,
                        randomized=True, max_steps=200000, lr_delay_steps=2500, lr_delay_mult=0.01, stage="radiance",
                        num_coarse_samples=NC, num_fine_samples=NF, near=NEAR, far=FAR)
    if a.train_batch % world:
        raise ValueError("train batch must divide by the GPU count (train.py:196)")
    B = a.train_batch // world
    rays_hw = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), 800, 800)
    flat = utils.namedtuple_map(lambda r: r.reshape(-1, r.shape[-1]), rays_hw)
    gen = torch.Generator().manual_seed(1234 + rank)
    idx = torch.randint(0, 640000, (B,), generator=gen)
    rays = utils.namedtuple_map(lambda r: r[idx].to(dev).contiguous(), flat)
    # the env patch is part of the batch dict the reference shards over its devices (datasets.py:98-100 -> utils.shard:
    # [128,128,.] -> [n_dev, 128/n_dev, 128, .]), so each rank evaluates 128/N rows of it (and train.py:128-129 reshapes its
    # shard to [ps, ps, -1] with ps = 128/N, which loss_fn reproduces)
    env = synthetic.blender_rays(synthetic.camera_pose(1.3, 0.8, 4.03), 128, 128, camera_angle_x=0.2)
    e0, e1 = utils.shard_range(128, rank, world)
    env = utils.namedtuple_map(lambda r: r[e0:e1].to(dev).contiguous(), env)
    batch = {"rays": rays, "pixels": torch.rand(B, 3, generator=gen).to(dev), "env_rays": env, "annealed_alpha": 0.5}
    out = {}
    for label, ws in (("with_allreduce", world),) + ((("compute_only", 1),) if world > 1 else ()):
        variables = model.init(0)
        state = train.TrainState.create(variables, targs, model=model)
        state.step, rng = 3000, 0
        for _ in range(max(4, a.warmup)):          # 2 eager steps, capture, replays
            state, stats, rng = train.train_step(model, rng, state, batch, targs, world_size=ws)
        barrier()
        l0 = _lib.launch_count() + state.replayed_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.train_steps):
            state, stats, rng = train.train_step(model, rng, state, batch, targs, world_size=ws)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        out[label] = (ms / a.train_steps, int(_lib.launch_count() + state.replayed_kernel_launches() - l0), float(stats["loss"]))
        torch.cuda.synchronize()
        state.graphs.clear()
        del state, variables
    ms, launches, loss = out["with_allreduce"]
    res = {"metric": "training rays/sec (fwd+bwd+allreduce+Adam)", "value": a.train_batch / (ms * 1e-3), "unit": "rays/s",
           "ms_per_step": ms, "global_batch": a.train_batch, "rays_per_gpu": B, "scaling": "strong", "steps": a.train_steps,
           "stage": "radiance", "cuda_graph": True, "gpu_launches": launches, "loss": loss,
           "allreduce": "none (N=1)" if world == 1 else
           "in-graph NCCL all-reduce(sum) of the fine/coarse/bkgd gradient buckets (4.99 MB fp32) + stats, then x1/N",
           "flop_per_ray": 3 * 256 * FLOP_PER_SAMPLE}
    tf = res["flop_per_ray"] * a.train_batch / (ms * 1e-3) / 1e12
    res["tflops_per_gpu"] = tf / world
    res["frac_of_bf16_sustained"] = tf / world / float(peaks.get("bf16_tflops_sustained", 1400.0))
    if world > 1:
        cms = out["compute_only"][0]
        res["compute_only_ms"] = cms
        res["exposed_comm_ms"] = ms - cms
        res["limiter"] = ("exposed all-reduce" if ms - cms > 0.5 * cms else
                          "per-rank latency that does not shrink with B/N: one-wave weight-gradient launches (TMEM allocation, "
                          "pipeline fill, 256 KB accumulator flush), the serial 768-step march, single-block chains of the "
                          "background MLP, ~60 launches of a few microseconds -- DESIGN.md section 6")
    return res


def main():
    a = parse()
    rank, world, local = dist_env()
    if a.impl == "reference":
        return run_reference(a, rank, world)

    from samplenerfro_b200 import _lib, models, ops, synthetic, utils
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    # ---- scene (setup, untimed): 512^3 ship-sized blob, blur 9/sigma 3 (configs/ship_*.gin), random-init weights
    G = a.grid
    ndim, nmin, nmax = [G] * 3, [-1.5] * 3, [1.5] * 3
    data = synthetic.ellipsoid_occupancy(G, 1.5, (1.0, 0.4, 0.6), ss=4, device=dev)
    n = ops.grid_blur(synthetic.rescale_ior(data, "ship_skydome-bkgd_no-partial-reflect_cycles"), ndim, 9, 3.0)
    del data
    flags = utils.Flags(config="ship_skydome-bkgd_no-partial-reflect_cycles", num_path_samples=P, white_bkgd=False,
                        use_online_sparsity=False, num_coarse_samples=NC, num_fine_samples=NF, near=NEAR, far=FAR)
    model, variables = models.construct_nerf(0, None, flags, ndim, nmin, nmax, n)
    del n
    H = W = a.side
    n_total = H * W
    # ONE frame, the same camera on every rank.  N > 1: the frame's rows are cut into N bands (strong scaling).
    rays_hw = synthetic.blender_rays(synthetic.camera_pose(0.7, 1.0, 4.03), H, W)
    host_hw = utils.namedtuple_map(lambda r: r.contiguous().pin_memory(), rays_hw)                     # [H,W,.] pinned
    dev_hw = utils.namedtuple_map(lambda r: r.to(dev, non_blocking=True), host_hw)                    # [H,W,.] resident
    dev_rays = utils.namedtuple_map(lambda r: r.reshape(n_total, -1), dev_hw)
    r0, r1, per_band = utils.band_rows(H, rank, world)
    torch.cuda.synchronize()

    chunk = a.chunk
    mlp_ev = []          # (start, stop, samples) CUDA events around every enc+MLP launch of the timed region
    record = {"on": False}
    orig_fwd = ops.encmlp_fwd

    def timed_encmlp(packed, pos, dirs, debug_layers=False):
        if not record["on"]:
            return orig_fwd(packed, pos, dirs, debug_layers)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_fwd(packed, pos, dirs, debug_layers)
        e1.record()
        mlp_ev.append((e0, e1, pos.numel() // 3))
        return out

    ops.encmlp_fwd = timed_encmlp

    # the memory-bound stages, timed the same way (CUDA events around their launches inside the timed region) and
    # reported against the measured HBM copy peak with the algorithmic bytes of SURVEY 8(d)
    stage_ev = {"march": [], "composite": [], "resample": []}
    orig_march, orig_comp, orig_resample = ops.march, ops.composite_fwd, ops.resample

    def _timed(name, fn, nbytes):
        def wrapper(*args, **kw):
            if not record["on"]:
                return fn(*args, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*args, **kw)
            e1.record()
            stage_ev[name].append((e0, e1, nbytes(args, kw)))
            return out
        return wrapper

    ops.march = _timed("march", orig_march, lambda a_, k_: a_[4].shape[0] * (24 + 44 * a_[8]))                 # origins, n_steps
    ops.composite_fwd = _timed("composite", orig_comp, lambda a_, k_: a_[1].shape[0] * (32 * a_[1].shape[1] + 36))   # t [B,Ns]
    ops.resample = _timed("resample", orig_resample, lambda a_, k_: a_[1].shape[0] * ((63 + 62) * 4 + 192 * 80))     # t_c [B,Nc]

    def apply_dev(k0, k1, r):
        return model.apply(variables, k0, k1, r, False)

    def apply_host(k0, k1, r):
        r = utils.namedtuple_map(lambda x: x.to(dev, non_blocking=True), r)
        return model.apply(variables, k0, k1, r, False)

    def render_resident():
        """One frame with the rays already in HBM; outputs stay on the device.  N > 1: this rank's band of the frame, then
        the NCCL all-gather of the bands (every rank ends up with the whole frame)."""
        if world == 1:
            outs = []
            for i in range(0, n_total, chunk):
                r = utils.namedtuple_map(lambda x: x[i:i + chunk], dev_rays)
                ret, _ = model.apply(variables, 1, 2, r, False)
                outs.append(ret[-1][0])
            return outs
        return utils.render_image_sharded(apply_dev, dev_hw, 0, False, chunk=chunk, rank=rank, world_size=world)

    out_host = torch.empty(H, W, 5, pin_memory=True)

    def render_e2e():
        """The public call a user makes: render_image over HOST rays; H2D of the rays and D2H of rgb/dist/acc inside.
        N > 1: render_image_sharded -- every rank uploads and renders its band, all-gather, rank 0 reads the frame back."""
        if world == 1:
            rgb, dist_, acc = utils.render_image(apply_host, host_hw, 0, False, chunk=chunk)
        else:
            rgb, dist_, acc = utils.render_image_sharded(apply_host, host_hw, 0, False, chunk=chunk, rank=rank,
                                                         world_size=world)
        if rank == 0:      # one contiguous [H,W,5] device tensor -> ONE device-to-host copy into pinned memory
            out_host.copy_(torch.cat([rgb, dist_, acc], dim=-1), non_blocking=True)
        torch.cuda.synchronize()
        return out_host

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, with_events=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        l0 = _lib.launch_count()
        record["on"] = with_events
        if with_events:      # `ncu --profile-from-start off` captures exactly the timed region (no-op otherwise)
            torch.cuda.cudart().cudaProfilerStart()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        if with_events:
            torch.cuda.cudart().cudaProfilerStop()
        record["on"] = False
        sampler.stop_flag = True
        sampler.join()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, _lib.launch_count() - l0, sampler.summary()

    ms, launches, clocks = timed(render_resident, a.steps, a.warmup, with_events=True)
    value = n_total * a.steps / (ms * 1e-3)
    # dominant kernel: the enc+MLP launches (92 % of the step)
    mlp_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in mlp_ev)
    mlp_samples = sum(m for _, _, m in mlp_ev)
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    ach_tf = FLOP_PER_SAMPLE * mlp_samples / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    fine_samples = max((m for _, _, m in mlp_ev), default=0)
    roofline = {"kernel": "encmlp_pair_kernel (pos_enc + NerfMLP, tcgen05 cta_group::2)", "bound": "tensor", "achieved": ach_tf,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)",
                "share_of_step": mlp_ms / ms, "launches": len(mlp_ev),
                "flop_per_launch_avg": FLOP_PER_SAMPLE * mlp_samples / max(1, len(mlp_ev)),
                "traffic": MLP_DRAM_BYTES.get(fine_samples),
                "traffic_unit": "dram bytes read+written by the largest (fine-pass) launch, ncu --set full at this launch size "
                                "(profiles/, see MLP_DRAM_BYTES in bench.py); null = no capture at this launch size"}
    hbm = float(peaks.get("hbm_gbs", 6500.0))
    stage_roof = []
    for name, evs in stage_ev.items():
        t_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in evs)
        nb = sum(b for _, _, b in evs)
        if t_ms > 0:
            stage_roof.append({"kernel": name, "bound": "hbm", "achieved": nb / (t_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                               "frac": nb / (t_ms * 1e-3) / 1e9 / hbm, "share_of_step": t_ms / ms, "launches": len(evs)})
    e2e_ms, _, _ = timed(render_e2e, a.steps, max(1, a.warmup - 1))
    h2d = sum(t.numel() * 4 for t in host_hw)          # over all ranks: every band is uploaded exactly once
    d2h = out_host.numel() * 4                        # rank 0 reads the assembled frame back
    e2e = {"value": n_total * a.steps / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.steps}
    if world == 1:
        par = "one GPU, whole frame per model.apply, no collective"
        coll = None
    else:
        par = (f"ONE frame row-sharded x{world} ({per_band} rows = {per_band * W} rays per rank), bands all-gathered over NCCL "
               "inside the timed region (eval.py:96-97)")
        coll = {"op": "all_gather", "bytes_per_rank": per_band * W * 5 * 4, "where": "utils.render_image_sharded"}
    line = {"metric": "rays/sec (march+MLP+composite)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16 MLP (fp32 accumulate) / f32 march+composite",
            "data": "synthetic",
            "config": {"workload": f"ship_skydome {H}x{W} refractive render: S={S} eikonal steps, IoR grid {G}^3, "
                                   f"{NC} coarse + {NC + NF} fine MLP samples/ray, random-init weights",
                       "rays_per_step": n_total, "rays_per_step_per_gpu": (r1 - r0) * W, "chunk": chunk,
                       "l2": f"inputs larger than L2: path {min(chunk, (r1 - r0) * W) * S * 36 / 2**20:.0f} MiB/chunk, table {G**3 * 16 / 2**20:.0f} MiB",
                       "parallelism": par},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "other_kernels": stage_roof}
    if coll is not None:
        line["collective"] = coll
    ops.encmlp_fwd, ops.march, ops.composite_fwd, ops.resample = orig_fwd, orig_march, orig_comp, orig_resample
    if not a.no_train:
        torch.cuda.empty_cache()
        line["train_step"] = train_arm(a, model, rank, world, local, dev, barrier, peaks)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_baseline(64)
        v, nr, dt = cpu_baseline(a.cpu_rays)
        line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{nr} rays of the same camera through the oracle port (torch-CPU fp32) in {dt:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        from samplenerfro_b200 import train
        train.shutdown_distributed(None)


if __name__ == "__main__":
    main()
