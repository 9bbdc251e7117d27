"""Turn the ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python scripts/summarize_profiles.py         # needs gpurun_out/launches_r1.csv and prof_k3_r1.ncu-rep
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
ROUND = "r1"

rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{ROUND}.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki].split("(")[0].replace("void ", "").split("<")[0], float(r[vi])) for r in rows[1:]]
ppn = [(n, v) for n, v in seq if n.startswith("ppn::")]
per_step = 2 if any("parse_fused" in n for n, _ in ppn) else 3
steps, chunks = ppn[:12 * per_step], ppn[12 * per_step:]    # 12 device steps, then the host-buffer arm's chunks


def table(lst, title):
    agg = collections.defaultdict(list)
    for n, v in lst:
        agg[n].append(v)
    tot = sum(sum(v) for v in agg.values())
    out = [title, f"{'kernel':42s} {'launches':>8s} {'median us':>10s} {'min us':>8s} {'share':>7s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        v2 = sorted(v)
        out.append(f"{k:42s} {len(v):8d} {v2[len(v2) // 2] / 1e3:10.2f} {v2[0] / 1e3:8.2f} {100 * sum(v) / tot:6.1f}%")
    return out


out = ["# ncu launch list of `python bench.py --steps 4 --warmup 3 --settle-s 0 --no-cpu-baseline --e2e-steps 1`  (B200, round 1)",
       "# metric gpu__time_duration.sum, --clock-control none.  Under ncu every launch is serialised and cold-cache:",
       "# compare the SHARES with bench.py's roofline.stage_ms_per_step (event-bracketed serial pass: K3 63.0 / K124 34.0 us -> 65 / 35 %),",
       "# not the absolutes.  In the timed region itself the two kernels of consecutive steps overlap (step = 56.5 us).",
       "# Raw list: profiles/launches_r1.csv", ""]
out += table(steps, "## device steps: 12 x ppn_parse on 512 images (warm-up 3+1, timed 4, per-kernel event pass 4): arg-max + fused parse")
out += [""] + table(chunks, "## end-to-end arm: 3 x ppn_parse_host = 24 chunks of 64 images (host buffers, copies overlapped)")
open(os.path.join(P, f"launches_{ROUND}_summary.txt"), "w").write("\n".join(out) + "\n")
shutil.copy(os.path.join(G, f"launches_{ROUND}.csv"), os.path.join(P, f"launches_{ROUND}.csv"))
print("\n".join(out))

rep = os.path.join(G, f"prof_k3_{ROUND}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput",
        "gpu__dram_throughput", "dram__cycles_active", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct",
        "smsp__issue_active.avg.pct", "lts__t_sector_hit_rate", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active",
        "launch__occupancy_limit"]
keep = [h for h in hdr if any(s in h for s in want)]
idx = [hdr.index(k) for k in keep]
with open(os.path.join(P, f"k3_ncu_full_{ROUND}.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(keep)
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
for r in rows[2:]:
    d = {hdr[i]: r[i] for i in idx}
    print({k: d[k] for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size")})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:limb_argmax"],
                     capture_output=True, text=True).stdout
open("/tmp/src_k3.csv", "w").write(src)
hot = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot.py"), "/tmp/src_k3.csv", "14"],
                     capture_output=True, text=True).stdout
open(os.path.join(P, f"k3_ncu_stalls_{ROUND}.txt"), "w").write(
    "# limb arg-max kernel (cfg2, 512 images): warp-stall samples per SASS instruction, from `ncu --set full --import-source on`\n" + hot)
print(hot[:600])

rep2 = os.path.join(G, f"prof_k124_{ROUND}.ncu-rep")
if os.path.exists(rep2):
    raw = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = [h for h in hdr if any(s in h for s in want)]
    idx = [hdr.index(k) for k in keep]
    with open(os.path.join(P, f"k124_ncu_full_{ROUND}.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    src = subprocess.run(["ncu", "-i", rep2, "--page", "source", "--csv", "--kernel-name", "regex:parse_fused"],
                         capture_output=True, text=True).stdout
    open("/tmp/src_k124.csv", "w").write(src)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot.py"), "/tmp/src_k124.csv", "20"],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f"k124_ncu_stalls_{ROUND}.txt"), "w").write(
        "# fused parse kernel (cfg2, 512 images): warp-stall samples per SASS instruction, from `ncu --set full --import-source on`\n" + hot)
    print(hot[:1500])
