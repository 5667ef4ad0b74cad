#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel launch) into the handful of counters DESIGN.md quotes.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json out.json]
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__maximum_warps_per_active_cycle_pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def unit_scale(u: str) -> float:
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, u = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")]}
        for k in KEYS:
            if k in h:
                i = h.index(k)
                try:
                    d[k] = float(r[i].replace(",", "")) * (unit_scale(u[i]) if ("bytes" in k or "time" in k) else 1.0)
                except ValueError:
                    d[k] = r[i]
        if "dram__bytes_read.sum" in d:
            d["dram_bytes_per_launch"] = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        out.append(d)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out[0] if len(out) == 1 else out, f, indent=1)
    for d in out:
        for k, v in d.items():
            print(f"{k:80s} {v}")


if __name__ == "__main__":
    main()
