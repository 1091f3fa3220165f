"""Extract the key metrics of every kernel in an .ncu-rep (ncu --set full capture) as label,metric,value,unit rows: the summaries
under profiles/.  Usage: python scripts/ncu_key_metrics.py report.ncu-rep "label for kernel 1" ["label for kernel 2" ...]"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
rep, labels = sys.argv[1], sys.argv[2:]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
for n, r in enumerate(rows[2:]):
    label = labels[n] if n < len(labels) else r[hdr.index("Kernel Name")]
    for m in METRICS:
        if m in hdr:
            w.writerow([label, m, r[hdr.index(m)], units[hdr.index(m)]])
