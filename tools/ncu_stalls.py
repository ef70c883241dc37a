#!/usr/bin/env python
"""Per-instruction stall summary from an .ncu-rep source page:  python tools/ncu_stalls.py rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [(j, n) for j, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
out = []; tot = 0; agg = {}
for r in rows[h + 1:]:
    if len(r) <= si or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name": break
        continue
    s = int(r[si] or 0); tot += s
    for j, n in stall_cols:
        agg[n] = agg.get(n, 0) + int(r[j] or 0)
    out.append((s, r[src].strip(), r[ie], sorted([(int(r[j] or 0), n) for j, n in stall_cols], reverse=True)[:2]))
print("total samples", tot, "instructions", len(out), "warp-instructions executed", sum(int(o[2] or 0) for o in out))
print("by reason:", ", ".join(f"{n[6:]} {100 * v / max(tot,1):.1f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for k, (s, ins, ie_, top) in enumerate(out):
    if s > tot * minpct / 100:
        print(f"{k:4d} {100 * s / tot:5.1f}%  {ins[:58]:58s} exec={ie_:>8s} {top}")
