"""Reduce a multi-result `ncu --set full` report (or its `--page raw --csv` dump) to the tracked text summary:
  python scripts/ncu_multi_summary.py gpurun_out/r3a/render_raw.csv profiles/r3a_render_ncu_summary.txt "<header note>"
One column block per captured launch, the metrics the DESIGN / VERDICT discussion uses."""
import csv, sys

src, dst = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
with open(dst, "w") as f:
    f.write(f"# {note}\n# ncu --set full --import-source on --clock-control none; one block per captured launch.  Never a bench value: ncu replays and serialises.\n")
    for r in data:
        f.write(f"\n== {r[idx['Kernel Name']][:140]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
        for k in KEYS:
            if k in idx:
                f.write(f"{k} [{units[idx[k]]}] = {r[idx[k]]}\n")
        try:
            rd = float(r[idx['dram__bytes_read.sum']].replace(",", "")); wr = float(r[idx['dram__bytes_write.sum']].replace(",", ""))
            ur, uw = units[idx['dram__bytes_read.sum']], units[idx['dram__bytes_write.sum']]
            sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            f.write(f"dram_total_bytes = {rd * sc[ur] + wr * sc[uw]:.0f}\n")
        except Exception as e:
            f.write(f"dram_total_bytes = ? ({e})\n")
print(open(dst).read())
