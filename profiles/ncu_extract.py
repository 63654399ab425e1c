"""Print the metrics we track from an `ncu --page raw --csv` export (one kernel per row)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__inst_executed_op_shfl", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"]
for r in rows[2:]:
    print("-" * 60)
    for h, u, v in zip(hdr, units, r):
        if any(h == k or (k in h and ("issue_stalled" in k or "shfl" in k)) for k in KEYS):
            if "issue_stalled" in h and "_per_warp_active.pct" not in h:
                continue
            print(f"{h} [{u}] = {v}")
