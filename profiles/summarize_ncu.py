#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into a small text summary for profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# source: {path}  (ncu --set full --clock-control none; cold-cache replay)")
    for n, r in enumerate(data):
        print(f"\n## launch {n}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:80s} {r[i]} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_")[0]))
                except ValueError:
                    pass
        if stalls:
            print("stall cycles per issued instruction: " +
                  ", ".join(f"{nm}={v:.2f}" for v, nm in sorted(stalls, reverse=True)[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
