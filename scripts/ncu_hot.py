#!/usr/bin/env python
"""Top stall lines of one kernel from an ncu report's source page (SASS view):
    python scripts/ncu_hot.py gpurun_out/x.ncu-rep rs_onesweep [topN]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[col["# Samples"]].isdigit()]
tot = sum(int(r[col["# Samples"]]) for r in body)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[col[h]]) for r in body) for h in stall_cols}
print(rows[0][1][:120])
print("total samples", tot, "| by reason:", ", ".join(f"{k[6:]} {v*100//max(tot,1)}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for idx, r in sorted(enumerate(body), key=lambda x: -int(x[1][col["# Samples"]]))[:top]:
    n = int(r[col["# Samples"]])
    why = sorted(((int(r[col[h]]), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{n*100/max(tot,1):5.1f}%  #{idx:4d} {r[col['Source']].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
