#!/usr/bin/env python3
"""ncu_src_top.py <source-page csv> [n] -- top stall sites of an `ncu --page source --csv` export, with the
dominant stall reason per instruction and the cumulative share of samples, plus totals per stall reason."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]; body = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in body)
print("kernel:", rows[0][1]); print("instructions:", len(body), " samples:", tot)
by_reason = {h: sum(int(r[ix[h]]) for r in body) for h in reasons}
print("by reason:", ", ".join(f"{h[6:]} {100*v/tot:.1f}%" for h, v in sorted(by_reason.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
top = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:n]
for i in sorted(top):
    r = body[i]; s = int(r[ix["# Samples"]])
    dom = max(reasons, key=lambda h: int(r[ix[h]]))
    print(f"{i:5d} {100*s/tot:5.2f}%  {dom[6:]:14s} {r[ix['Source']].strip()[:90]}")
