#!/usr/bin/env python
"""ncu_summary.py <report.ncu-rep> <out.txt> [comment ...] — selected counters of every profiled launch in a report, one "name value unit"
line each (the format bench.py's ncu_traffic() reads), for the tracked summaries under profiles/."""
import csv, subprocess, sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum", "sm__inst_executed_pipe_tensor_subpipe_imma.sum",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores")
KEEP_SUB = ("issue_stalled", "pipe_tensor")

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    for c in sys.argv[3:]:
        f.write(f"# {c}\n")
    for r in rows[2:]:
        f.write(f"{'Kernel Name':<80} {r[hdr.index('Kernel Name')]}\n{'Block Size':<80} {r[hdr.index('Block Size')]}\n{'Grid Size':<80} {r[hdr.index('Grid Size')]}\n")
        for i, h in enumerate(hdr):
            if h in KEEP or (any(s in h for s in KEEP_SUB) and ("per_issue_active.ratio" in h or "pct_of_peak_sustained_active" in h)):
                f.write(f"{h:<80} {r[i]} {units[i]}\n")
        f.write("\n")
print(out)
