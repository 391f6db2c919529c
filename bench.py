#!/usr/bin/env python
"""Headline benchmark: rays/s of the refractive rendering hot path (march + enc/MLP + composite).

Workload (BASELINE.json configs[1]): ship_skydome synthetic scene, full 800x800 refractive render, S=768
march steps through a 512^3 IoR grid, 64 coarse + 192 fine radiance-MLP evaluations per ray, random-init
weights, synthetic blob -- one "step" is one full frame (640 000 rays).

  python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                   # CPU baseline (oracle port of the JAX path)

Prints ONE JSON line (see DESIGN.md "Measurement").  Under torchrun each rank renders a contiguous band of the
frame (no data-path collective); rank 0 reports total rays / max-over-ranks device time.
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
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ship_skydome 800x800 refractive render, S=768, 64 coarse + 192 fine samples/ray",
                       "sample": f"{n} rays per step of the same camera/pipeline on CPU"},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{n} rays/step x {a.steps} steps, oracle port (torch-CPU fp32) of the JAX path"},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ CUDA arm
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
    # weak scaling: per-GPU work is fixed -- every rank renders its OWN full frame (a different camera azimuth), so
    # N GPUs render N frames per step with no data-path collective.
    rays_hw = synthetic.blender_rays(synthetic.camera_pose(0.7 + 0.37 * rank, 1.0, 4.03), H, W)
    host = utils.namedtuple_map(lambda r: r.reshape(n_total, -1).contiguous().pin_memory(), rays_hw)
    dev_rays = utils.namedtuple_map(lambda r: r.to(dev, non_blocking=True), host)
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

    def render_resident():
        """One frame with the rays already in HBM; outputs stay on the device."""
        outs = []
        for i in range(0, n_total, chunk):
            r = utils.namedtuple_map(lambda x: x[i:i + chunk], dev_rays)
            ret, _ = model.apply(variables, 1, 2, r, False)
            outs.append(ret[-1][0])
        return outs

    out_host = torch.empty(H, W, 5, pin_memory=True)

    def render_e2e():
        """The public call a user makes: render_image over HOST rays; H2D of the rays and D2H of rgb/dist/acc inside."""
        def fn(k0, k1, r):
            r = utils.namedtuple_map(lambda x: x.to(dev, non_blocking=True), r)
            return model.apply(variables, k0, k1, r, False)
        rgb, dist_, acc = utils.render_image(fn, utils.namedtuple_map(lambda r: r.reshape(H, W, -1), host), 0, False,
                                             chunk=chunk)
        out_host[..., 0:3].copy_(rgb, non_blocking=True)
        out_host[..., 3:4].copy_(dist_, non_blocking=True)
        out_host[..., 4:5].copy_(acc, non_blocking=True)
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
    value = world * n_total * a.steps / (ms * 1e-3)
    # dominant kernel: the enc+MLP launches (92 % of the step)
    mlp_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in mlp_ev)
    mlp_samples = sum(m for _, _, m in mlp_ev)
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    ach_tf = FLOP_PER_SAMPLE * mlp_samples / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    roofline = {"kernel": "encmlp_pair_kernel (pos_enc + NerfMLP, tcgen05 cta_group::2)", "bound": "tensor", "achieved": ach_tf,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)",
                "share_of_step": mlp_ms / ms, "launches": len(mlp_ev),
                "traffic": MLP_DRAM_BYTES_PER_SAMPLE * mlp_samples / max(1, len(mlp_ev)),
                "traffic_unit": "bytes per launch (ncu dram bytes per sample x samples per launch)"}
    hbm = float(peaks.get("hbm_gbs", 6500.0))
    stage_roof = []
    for name, evs in stage_ev.items():
        t_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in evs)
        nb = sum(b for _, _, b in evs)
        if t_ms > 0:
            stage_roof.append({"kernel": name, "bound": "hbm", "achieved": nb / (t_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                               "frac": nb / (t_ms * 1e-3) / 1e9 / hbm, "share_of_step": t_ms / ms, "launches": len(evs)})
    e2e_ms, _, _ = timed(render_e2e, a.steps, max(1, a.warmup - 1))
    h2d = sum(t.numel() * 4 for t in host)
    d2h = out_host.numel() * 4
    e2e = {"value": world * n_total * a.steps / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.steps}
    line = {"metric": "rays/sec (march+MLP+composite)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 MLP (fp32 accumulate) / f32 march+composite",
            "data": "synthetic",
            "config": {"workload": f"ship_skydome {H}x{W} refractive render: S={S} eikonal steps, IoR grid {G}^3, "
                                   f"{NC} coarse + {NC + NF} fine MLP samples/ray, random-init weights",
                       "rays_per_step_per_gpu": n_total, "chunk": chunk,
                       "l2": f"inputs larger than L2: path {chunk * S * 36 / 2**20:.0f} MiB/chunk, table {G**3 * 16 / 2**20:.0f} MiB",
                       "parallelism": f"ray-sharded x{world}, no data-path collective"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "other_kernels": stage_roof}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_baseline(64)
        v, nr, dt = cpu_baseline(a.cpu_rays)
        line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{nr} rays of the same camera through the oracle port (torch-CPU fp32) in {dt:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
