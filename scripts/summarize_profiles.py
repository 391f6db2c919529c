"""Turn a gpurun_out/<tag>/ visit into the tracked summaries under profiles/ (run on the CPU box).
  python scripts/summarize_profiles.py gpurun_out/r1a r1a
Writes profiles/<tag>_launches_summary.txt (+ the raw csv) and profiles/<tag>_<kernel>_ncu_summary.txt."""
import csv, glob, os, subprocess, sys
from collections import OrderedDict

src, tag = sys.argv[1], sys.argv[2]
os.makedirs("profiles", exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"]

lc = os.path.join(src, "launches_bench.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(l for l in open(lc) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + v)
    tot = sum(t for _, t in agg.values())
    with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
        f.write(f"# {tag} launch list of the timed region of `python bench.py --steps 1 --warmup 1 --no-cpu-baseline`\n"
                "# (ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:\n"
                "# compare SHARES with the bench line, not absolutes).  kernel, launches, total_ms, share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k}, {n}, {t:.3f}, {100 * t / tot:.1f}%\n")
        f.write(f"TOTAL, {sum(n for n, _ in agg.values())}, {tot:.3f}, 100%\n")
    with open(f"profiles/{tag}_launches_bench.csv", "w") as f:
        f.write(open(lc).read())
    print(open(f"profiles/{tag}_launches_summary.txt").read())

for rep in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
    kn = os.path.basename(rep)[:-8]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print("no data in", rep); continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    with open(f"profiles/{tag}_{kn}_ncu_summary.txt", "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none, {d.get('Kernel Name', ('?',))[0]}\n"
                f"# grid {d.get('Grid Size', ('?',))[0]} block {d.get('Block Size', ('?',))[0]}; launch taken from scripts/perf_probe.py --rays 65536 "
                "(ship workload: G=512, S=768).  Never a bench value: ncu replays and serialises.\n")
        for k in KEYS:
            if k in d:
                f.write(f"{k} [{d[k][1]}] = {d[k][0]}\n")
    print(open(f"profiles/{tag}_{kn}_ncu_summary.txt").read())
