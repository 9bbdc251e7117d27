"""Summarise an ncu --page source --csv dump: instructions with the most stall samples.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv
    python scripts/ncu_hot.py src.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = []
for r in rows[2:]:
    if len(r) != len(hdr):
        if body:
            break                      # next kernel section: keep only the first
        continue
    if r[col["# Samples"]].isdigit():
        body.append(r)
    elif body:
        break
tot = sum(int(r[col["# Samples"]]) for r in body)
print("total samples", tot, " instructions", len(body), " warp-instr executed", sum(int(r[col["Instructions Executed"]]) for r in body))
agg = {s: sum(int(r[col[s]]) for r in body) for s in stalls}
print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]]))[:top]
for i in sorted(order):
    r = body[i]
    why = {s[6:]: int(r[col[s]]) for s in stalls if int(r[col[s]])}
    print(f"{i:5d} {int(r[col['# Samples']]):6d} exec={r[col['Instructions Executed']]:>8} {r[col['Source']].strip()[:70]:70s} {why}")
