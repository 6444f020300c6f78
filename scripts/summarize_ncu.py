#!/usr/bin/env python3
"""Summarise an `ncu --set full` capture (.ncu-rep) into a small text file for profiles/ and record the measured DRAM
traffic per launch in profiles/traffic.json (read by bench.py for roofline.traffic).

usage: summarize_ncu.py <report.ncu-rep> <out.txt> [traffic_key]     e.g. traffic_key = C2_f32
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    raw = page(rep, "raw")
    h, u, r = raw[0], raw[1], raw[2]
    col = {k: i for i, k in enumerate(h)}
    lines = [f"# {os.path.basename(rep)}: ncu --set full --clock-control none --import-source on (one launch)",
             f"kernel: {r[col['Kernel Name']]}"]
    for k in KEYS:
        if k in col:
            lines.append(f"{k} [{u[col[k]]}] = {r[col[k]]}")
    if "dram__bytes_read.sum" in col:
        rd = float(r[col["dram__bytes_read.sum"]]) * UNIT[u[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * UNIT[u[col["dram__bytes_write.sum"]]]
        lines.append(f"dram traffic per launch (read + write) = {rd + wr:.6g} bytes")
        if key:
            tpath = os.path.join(os.path.dirname(os.path.abspath(dst)), "traffic.json")
            t = json.load(open(tpath)) if os.path.exists(tpath) else {}
            t[key] = rd + wr
            json.dump(t, open(tpath, "w"), indent=1, sort_keys=True)
    stalls = sorted(((int(float(r[i])), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, i in col.items()
                     if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")), reverse=True)
    tot = sum(s for s, _ in stalls) or 1
    lines.append("warp-state samples: " + ", ".join(f"{n} {100.0 * s / tot:.1f}%" for s, n in stalls if s * 50 > tot))
    src = page(rep, "source")
    if len(src) > 2:
        sh = src[1]
        ia, isrc, ins, iex = sh.index("Address"), sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
        data = [x for x in src[2:] if len(x) > ins and x[ins].isdigit()]
        total = sum(int(x[ins]) for x in data) or 1
        base = int(data[0][ia], 16)
        lines.append(f"top instructions by samples (of {total}):  offset  samples  executed  SASS")
        for x in sorted(data, key=lambda x: -int(x[ins]))[:25]:
            lines.append(f"  {int(x[ia], 16) - base:#07x} {int(x[ins]):7d} {x[iex]:>10s}  {x[isrc].strip()[:90]}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:24]))


if __name__ == "__main__":
    main()
