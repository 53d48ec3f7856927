#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the JSON committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/rNN_ncu_full_x.json

Reads the report with `ncu -i REP --page raw --csv` (run where ncu is installed; no GPU needed) and keeps, per
kernel launch, the counters DESIGN.md argues from: duration, DRAM bytes, pipe utilisation, occupancy limits,
shared-memory wavefronts (actual vs ideal) and the top stall reasons per issued instruction."""
import csv
import json
import subprocess
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal", "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):  # already the raw page (`ncu -i REP --page raw --csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {}
        for k in KEEP:
            if k in d and d[k] != "":
                rec[k] = (d[k] + (" " + u[k] if u.get(k) else "")).strip()
        stalls = {}
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "_not_issued" not in h:
                try:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(d[h].replace(",", "")), 3)
                except ValueError:
                    pass
        rec["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        out.append(rec)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
