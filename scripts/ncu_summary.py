"""Condense an .ncu-rep (raw page) into the metrics the round notes cite: python scripts/ncu_summary.py in.ncu-rep out.csv"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration", "dram__bytes", "dram__throughput", "gpu__dram_throughput", "l1tex__data_pipe_lsu_wavefronts", "lts__t_sector_hit_rate",
        "lts__throughput", "l1tex__throughput", "sm__throughput", "pipe_alu", "pipe_lsu", "inst_executed.sum", "registers_per_thread", "issue_active",
        "issue_stalled", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "sm__warps_active", "launch__waves", "launch__occupancy",
        "shared_mem_per_block", "pipe_tensor"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "metric", "unit", "value"])
    for li, vals in enumerate(rows[2:]):
        for i, h in enumerate(hdr):
            if any(k in h for k in KEEP) and vals[i] != "":
                w.writerow([li, h, units[i], vals[i]])
