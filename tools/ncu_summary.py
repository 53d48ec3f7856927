#!/usr/bin/env python3
"""Condense `ncu -i report --page raw --csv` into the JSON summaries kept under profiles/: one record per captured
launch with the metrics DESIGN.md and bench.py's `roofline.traffic` refer to (values keep ncu's units).
usage: tools/ncu_summary.py raw.csv out.json"""
import csv
import json
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active",
]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {}
        for k in KEEP:
            if k in d and d[k] not in ("", "n/a"):
                rec[k] = d[k] if k in ("Kernel Name", "Grid Size", "Block Size") or not unit.get(k) else f"{d[k]} {unit[k]}"
        stalls = {}
        for k, v in d.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a"):
                stalls[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(v.replace(",", "")), 3)
        rec["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        out.append(rec)
    json.dump(out, open(dst, "w"), indent=1)
    for rec in out:
        print(rec.get("Kernel Name", "?")[:70], rec.get("gpu__time_duration.sum"), rec.get("dram__bytes_read.sum"), rec.get("dram__bytes_write.sum"),
              rec.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
