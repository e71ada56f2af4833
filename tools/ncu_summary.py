#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into the handful of numbers DESIGN.md / bench.py quote.
Usage: python tools/ncu_summary.py gpurun_out/r01_embed.ncu-rep > profiles/r01_embed_ncu_summary.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = {"report": rep, "kernels": []}
    for d in data:
        k = {"name": d[hdr.index("Kernel Name")]}
        stalls = {}
        for i, h in enumerate(hdr):
            if h in KEYS:
                k[h] = {"value": d[i], "unit": units[i]}
            elif h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
                stalls[h[len(STALLS):-len("_per_issue_active.ratio")]] = float(d[i].replace(",", ""))
        k["warps_stalled_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        res["kernels"].append(k)
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
