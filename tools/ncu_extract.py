#!/usr/bin/env python
"""Pull the handful of counters the roofline argument needs out of an `ncu --page raw --csv` export.

    python tools/ncu_extract.py profiles/raw/x_raw.csv [--json out.json]
Prints one markdown table row per profiled launch; --json writes {kernel: {traffic_bytes, time_us, ...}} averaged per
kernel name (bench.py reads profiles/traffic.json for roofline.traffic)."""
import csv
import json
import re
import sys
from collections import defaultdict

COLS = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__inst_executed.sum", "warp_insts"), ("lts__t_sector_hit_rate.pct", "l2hit%")]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def short(name):
    m = re.match(r"(?:void )?(?:bvb::)?(\w+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in COLS) + " | traffic_MB |")
    print("|---|" + "---|" * (len(COLS) + 1))
    agg = defaultdict(lambda: defaultdict(list))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        vals = {}
        for m, n in COLS:
            if m not in idx:
                vals[n] = None
                continue
            v = float(r[idx[m]].replace(",", "") or "nan") * SCALE.get(units[idx[m]], 1.0)
            vals[n] = v
        vals["traffic_MB"] = (vals["dram_rd_MB"] or 0) + (vals["dram_wr_MB"] or 0)
        k = short(r[idx["Kernel Name"]])
        print(f"| {k} | " + " | ".join("-" if vals[n] is None else f"{vals[n]:.4g}" for _, n in COLS) + f" | {vals['traffic_MB']:.1f} |")
        for n, v in vals.items():
            if v is not None:
                agg[k][n].append(v)
    if "--json" in sys.argv:
        out = {k: {n: sum(v) / len(v) for n, v in d.items()} for k, d in agg.items()}
        for k in out:
            out[k]["traffic_bytes"] = out[k]["traffic_MB"] * 1e6
            out[k]["launches_profiled"] = len(agg[k]["time_us"])
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
