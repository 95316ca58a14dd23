#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-launch CSV kept under profiles/ (the .ncu-rep itself is scratch).
usage: python profiles/ncu_summary.py REPORT.ncu-rep OUT.csv
Runs `ncu -i REPORT --page raw --csv` and keeps the columns that the roofline argument needs: duration, DRAM bytes,
DRAM / SM / tensor-pipe / FMA-pipe utilisation, L2 hit rate, occupancy, registers, grid, shared-memory bank conflicts."""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__cluster_size",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [k for k in KEEP if k in idx]
    # any tensor-pipe metric the report has (names differ between ncu versions)
    for h in hdr:
        if ("pipe_tensor" in h or "tmem" in h) and h not in cols and ("pct" in h or h.endswith(".sum")):
            cols.append(h)
    kname = idx.get("Kernel Name", idx.get("Function Name"))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel"] + [f"{c} [{units[idx[c]]}]" for c in cols])
        for r in body:
            w.writerow([r[idx["ID"]], r[kname][:110]] + [r[idx[c]] for c in cols])
    print(f"{out}: {len(body)} launches, {len(cols)} metrics")


if __name__ == "__main__":
    main()
