"""Per-instruction sample table of the hottest loop from `ncu -i X.ncu-rep --page source --csv` (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try:
        data.append((r[ix["Address"]], r[ix["Source"]], int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0),
                     {s: int(r[ix[s]] or 0) for s in stalls}))
    except ValueError:
        pass
tot = sum(d[2] for d in data)
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
print("total samples", tot)
agg = {}
for a, src, n, ex, st in data:
    for k, v in st.items(): agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join(f"{k[6:]} {100*v/max(1,sum(agg.values())):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
cum = 0
for a, src, n, ex, st in data:
    if ex < thr: continue
    cum += n
    top = max(st.items(), key=lambda kv: kv[1])
    print(f"{a[-5:]} {ex:8d} {n:6d} {100*n/tot:5.1f}% cum {100*cum/tot:5.1f}%  {top[0][6:]:>16s}  {src[:70]}")
