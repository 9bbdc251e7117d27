"""Warp-stall samples per SOURCE line of one kernel from an ncu report captured with --import-source on.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
out, fname, hdr = [], None, None
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) >= 8 and r[0].isdigit():
        try:
            samp, inst = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        if samp > 0:
            out.append((fname, int(r[0]), samp, inst, r[1].strip()[:110]))
tot = sum(o[2] for o in out)
print("total samples", tot, " warp-instructions", sum(o[3] for o in out))
for o in sorted(out, key=lambda x: -x[2])[:top]:
    print(f"{o[0]:15s} {o[1]:5d} {o[2]:5d} {100 * o[2] / tot:5.1f}% inst={o[3]:9d}  {o[4]}")
