#!/usr/bin/env python
"""Condense an .ncu-rep (read here, no GPU needed) into the per-kernel lines we cite:
duration, DRAM bytes, instruction counts, occupancy, top stall reasons."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    print(f"# {path}")
    for n, r in enumerate(rows[2:]):
        print(f"\n## launch {n}: {r[idx['Kernel Name']]}")
        for w in WANT:
            if w in idx:
                print(f"{w:72s} {r[idx[w]]} {units[idx[w]]}")
        st = sorted(((float(r[idx[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace(
            "_per_issue_active.ratio", "")) for h in stall), reverse=True)[:6]
        print("top stalls (warps per issue-active):", ", ".join(f"{k} {v:.2f}" for v, k in st))


if __name__ == "__main__":
    main(sys.argv[1])
