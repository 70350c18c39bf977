#!/usr/bin/env python3
"""Summarises `ncu --set full` captures (raw CSV pages) into profiles/<round>_traffic.json: per kernel
DRAM bytes per launch, duration under ncu, pipe utilisation, occupancy, registers, top stall reasons.

    ncu -i X.ncu-rep --page raw --csv > X.raw.csv
    python tools/ncu_summary.py out.json name=X.raw.csv[:row] ...
`row` picks the launch inside the capture (default 0)."""
import csv
import json
import sys

KEYS = {
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "duration_ms_under_ncu": "gpu__time_duration.sum",
    "fmaheavy_pipe_cycles_active_pct": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "fma_pipe_cycles_active_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "alu_pipe_cycles_active_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}


def main():
    out_path, specs = sys.argv[1], sys.argv[2:]
    out = {}
    for spec in specs:
        name, path = spec.split("=", 1)
        row = 0
        if ":" in path:
            path, row = path.rsplit(":", 1)
            row = int(row)
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        r = data[row]
        col = {h: i for i, h in enumerate(hdr)}
        rec = {"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for k, metric in KEYS.items():
            if metric in col:
                v = float(r[col[metric]].replace(",", ""))
                rec[k] = v * UNIT_SCALE.get(units[col[metric]], 1.0)
        rec["traffic"] = rec.get("dram_bytes_read", 0.0) + rec.get("dram_bytes_write", 0.0)
        stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(r[i]) for h, i in col.items()
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
        rec["stall_cycles_per_issue_top"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        out[name] = rec
    out["_source"] = "ncu --set full --clock-control none via tools/profile_target.py; raw CSV pages beside this file"
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
