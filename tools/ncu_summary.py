#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): headline metrics, stall mix, per-region samples."""
import collections, csv, subprocess, sys, io

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    print(d["Kernel Name"][:120], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    items = [(k, float(v.replace(",", ""))) for k, v in d.items()
             if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued") and v not in ("", "n/a")]
    tot = sum(v for _, v in items) or 1
    print("  stall mix: " + ", ".join(f"{k[33:]} {100 * v / tot:.1f}%" for k, v in sorted(items, key=lambda t: -t[1])[:9]))
