"""Text summary of one `ncu --set full` capture: python scripts/summarize_ncu.py X.ncu-rep > profiles/NAME.txt
Prints the launch header metrics, DRAM traffic of the launch, the warp-state mix and the hottest SASS instructions."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
print(f"# {rep.split('/')[-1]}: ncu --set full --clock-control none --import-source on (one launch)")
print("kernel:", vals[col["Kernel Name"]])
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
got = {}
for w in want:
    if w in col:
        print(f"{w} [{units[col[w]]}] = {vals[col[w]]}")
        got[w] = (float(vals[col[w]].replace(",", "")), units[col[w]])
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
try:
    tr = sum(got[k][0] * scale[got[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    print(f"dram traffic per launch (read + write) = {tr:.6g} bytes")
except Exception as e:
    print("dram traffic: n/a", e)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {x: i for i, x in enumerate(h)}
stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
data, agg = [], {}
for r in rows[2:]:
    if len(r) < len(h): continue
    try:
        n = int(r[ix["# Samples"]] or 0); ex = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    data.append((r[ix["Address"]], n, ex, r[ix["Source"]]))
    for s_ in stalls:
        agg[s_] = agg.get(s_, 0) + int(r[ix[s_]] or 0)
tot = sum(agg.values()) or 1
print("warp-state samples: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
nsamp = sum(d[1] for d in data)
print(f"top instructions by samples (of {nsamp}):  offset  samples  executed  SASS")
for a, n, ex, s_ in sorted(data, key=lambda d: -d[1])[:22]:
    print(f"  0x{a[-5:]} {n:7d} {ex:10d}  {s_[:70]}")
