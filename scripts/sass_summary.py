"""Mnemonic counts per kernel of the built library: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "lp_mp_b200/liblpmp_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UCGABAR", "CREDUX", "SHFL", "LDGSTS", "HMMA", "UTC", "LDTM"]
stats = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"<.*", "", d.replace("void ", "")).split("(")[0].split("::")[-1]
        cur = stats.setdefault(name, collections.Counter())
        cur["inst"] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur["SASS"] += 1
        for c in cols:
            if m.group(1).startswith(c):
                cur[c] += 1
print(f"# SASS summary of {lib} (cuobjdump -sass, sm_100a only; counts of mnemonics per kernel, all instantiations)")
print("# TMA = UTMALDG / UTMASTG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk), SYNCS (mbarrier), UCGABAR (cluster barrier), CREDUX (warp min reduction)")
print("# no tensor-core instructions anywhere (min-plus is not a contraction): HMMA / UTC*MMA / LDTM = 0\n")
print(f"{'kernel (all instantiations)':34s}" + "".join(f"{c:>8s}" for c in ["inst", "SASS"] + cols))
for name, c in stats.items():
    print(f"{name:34s}" + "".join(f"{c[k]:8d}" for k in ["inst", "SASS"] + cols))
