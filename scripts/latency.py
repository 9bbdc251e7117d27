"""Latency of ONE ppn_parse call on one image (BASELINE.json configs[0] and the reference's native shape, B = 1):
what rt_test.py:109-133 costs per frame once the head tensor is in HBM.

Three clocks, all on inputs that rotate over more than L2 so that every call reads a cold image:
  * events   CUDA events around one isolated call (synchronised before and after), median over the calls
  * device   %globaltimer span first-CTA-start -> last-CTA-end per kernel of that call (ppn_timeline): no launch gaps
  * stream   calls issued back to back without synchronising: the per-call period a frame loop sees

    python scripts/latency.py [--configs cfg2,native] [--calls 40] [--tune key=value ...]
"""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_pose_proposal_network_b200 import _lib  # noqa: E402
from pytorch_pose_proposal_network_b200.config import PRESETS  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="cfg2,native")
ap.add_argument("--calls", type=int, default=40)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--dist", default="U", choices=["U", "S"], help="SURVEY 8(d) input distribution: U uniform (a third to a half of the cells are person candidates), S sparse (resp = u^8, boxes x0.3: a few dozen candidates, few overlaps)")
ap.add_argument("--rotate-mb", type=int, default=160, help="inputs rotate over this many MB (more than L2 = every call cold)")
ap.add_argument("--tune", action="append", default=[])
args = ap.parse_args()
for kv in args.tune:
    k, v = kv.split("=")
    _lib.tune(**{k.replace(".", "_"): int(v)})

try:
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
    sm_mhz = lambda: pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
except Exception:                                               # noqa: BLE001 - no NVML: clocks simply not reported
    sm_mhz = lambda: -1

_a = torch.randn(8192, 8192, device="cuda")


def spin_up(ms=300):
    """Sustained load right before a measurement: isolated microsecond calls with host synchronisation between them
    leave the GPU idle most of the time, and an idle GPU drops its SM clock — a frame loop never sees that state
    (the network runs before the parse)."""
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    while True:
        for _ in range(4):
            _a @ _a
        t1.record(); t1.synchronize()
        if t0.elapsed_time(t1) > ms:
            return


for name in args.configs.split(","):
    cfg = PRESETS[name]()
    B = args.batch
    parser = PoseParser(cfg)
    img_bytes = B * cfg.C * cfg.HW * 4
    n_buf = max(2, -(-(args.rotate_mb << 20) // img_bytes))
    gen = torch.Generator(device="cuda").manual_seed(11)
    bufs = [torch.rand(B, cfg.C, cfg.H, cfg.W, device="cuda", generator=gen) for _ in range(n_buf)]
    if args.dist == "S":
        for t in bufs:
            r2 = t[:, :cfg.K] * t[:, :cfg.K]
            r4 = r2 * r2
            t[:, :cfg.K] = r4 * r4
            t[:, 4 * cfg.K:6 * cfg.K] *= 0.3
    out = parser.alloc_output(B)
    for i in range(10):
        parser.parse(bufs[i % n_buf], out=out)
    torch.cuda.synchronize()
    plan = parser.parse_plan(B)
    n_k = plan["launches"]

    spin_up()
    clk = [sm_mhz()]
    ev = []
    for i in range(args.calls):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        parser.parse(bufs[i % n_buf], out=out)
        b.record()
        torch.cuda.synchronize()
        ev.append(a.elapsed_time(b) * 1e3)

    clk.append(sm_mhz())
    spin_up()
    rec = torch.zeros(args.calls * n_k, 4, dtype=torch.int64, device="cuda")
    rec[:, 0] = rec[:, 2] = torch.iinfo(torch.int64).max
    _lib.check(_lib.lib().ppn_timeline(rec.data_ptr(), rec.shape[0]), "ppn_timeline")
    for i in range(args.calls):
        parser.parse(bufs[i % n_buf], out=out)
        torch.cuda.synchronize()
    _lib.check(_lib.lib().ppn_timeline(None, 0), "ppn_timeline")
    r = rec.cpu().numpy().reshape(args.calls, n_k, 4)
    span = [(int(c[:, 1].max()) - int(c[:, 0].min())) / 1e3 for c in r]
    per_k = [[(int(c[j, 1]) - int(c[j, 0])) / 1e3 for c in r] for j in range(n_k)]

    # where the parse kernel's time goes: one phase boundary per run (tune key timeline.phase), first CTA
    clk.append(sm_mhz())
    phases = {}
    for ph in list(range(1, 8)) + list(range(11, 20)):
        _lib.tune(timeline_phase=ph)
        spin_up(100)
        rec.zero_()
        rec[:, 0] = rec[:, 2] = rec[:, 3] = torch.iinfo(torch.int64).max
        _lib.check(_lib.lib().ppn_timeline(rec.data_ptr(), rec.shape[0]), "ppn_timeline")
        for i in range(args.calls):
            parser.parse(bufs[i % n_buf], out=out)
            torch.cuda.synchronize()
        _lib.check(_lib.lib().ppn_timeline(None, 0), "ppn_timeline")
        rr = rec.cpu().numpy().reshape(args.calls, n_k, 4)[:, n_k - 1]
        ok = rr[:, 3] < (1 << 62)
        if ok.any():
            phases[ph] = statistics.median(((rr[ok, 3] - rr[ok, 0]) / 1e3).tolist())
    k3 = {}
    for ph in (22, 23, 24, 25, 121, 122):                                  # the arg-max (cluster) kernel's own phases
        _lib.tune(timeline_phase=ph)
        spin_up(100)
        rec.zero_()
        rec[:, 0] = rec[:, 2] = rec[:, 3] = torch.iinfo(torch.int64).max
        if ph >= 100:
            rec[:, 3] = 0
        _lib.check(_lib.lib().ppn_timeline(rec.data_ptr(), rec.shape[0]), "ppn_timeline")
        for i in range(args.calls):
            parser.parse(bufs[i % n_buf], out=out)
            torch.cuda.synchronize()
        _lib.check(_lib.lib().ppn_timeline(None, 0), "ppn_timeline")
        rr = rec.cpu().numpy().reshape(args.calls, n_k, 4)[:, 0]
        ok = (rr[:, 3] < (1 << 62)) & (rr[:, 3] > 0)
        if ok.any():
            k3[ph] = statistics.median(((rr[ok, 3] - rr[ok, 0]) / 1e3).tolist())
    _lib.tune(timeline_phase=0)
    waited = statistics.median(((r[:, n_k - 1, 2] - r[:, n_k - 1, 0]) / 1e3).tolist())

    spin_up()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(args.calls * 5):
        parser.parse(bufs[i % n_buf], out=out)
    b.record()
    torch.cuda.synchronize()
    stream = a.elapsed_time(b) * 1e3 / (args.calls * 5)

    clk.append(sm_mhz())
    # the same call replayed from a CUDA graph on one buffer that is refilled from the rotating inputs (a device copy,
    # outside the events): what PoseParser.capture() gives a frame loop
    static = torch.empty_like(bufs[0])
    cap = parser.capture(static)
    spin_up()
    gr = []
    for i in range(args.calls):
        static.copy_(bufs[i % n_buf])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        cap.replay()
        b.record()
        torch.cuda.synchronize()
        gr.append(a.elapsed_time(b) * 1e3)
    import time
    wall = []
    for i in range(args.calls):
        static.copy_(bufs[i % n_buf])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cap.replay()
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e6)
    wall_plain = []
    for i in range(args.calls):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        parser.parse(bufs[i % n_buf], out=out)
        torch.cuda.synchronize()
        wall_plain.append((time.perf_counter() - t0) * 1e6)

    print(f"{name} B={B} dist {args.dist}: {img_bytes / 1e6:.1f} MB per call, {n_k} launches, {float(out.count.float().mean()):.0f} humans, tune={args.tune}, inputs rotate over {n_buf * img_bytes / 1e6:.0f} MB, "
          f"SM clock sampled around the loops {clk} MHz")
    print(f"   events around one isolated call   median {statistics.median(ev):6.1f} us   min {min(ev):6.1f} us")
    print(f"   events around one graph replay    median {statistics.median(gr):6.1f} us   min {min(gr):6.1f} us   (the buffer is L2-warm: just copied)")
    print(f"   host wall, call -> synchronised   median {statistics.median(wall_plain):6.1f} us plain, {statistics.median(wall):6.1f} us graph replay")
    print(f"   device span (first start->last end) median {statistics.median(span):6.1f} us   min {min(span):6.1f} us   "
          + "  ".join(f"k{j} {statistics.median(per_k[j]):.1f}" for j in range(n_k)))
    print("   parse kernel, us after its start: " + "  ".join(f"p{k} {v:.1f}" for k, v in phases.items()) + f"  past-wait {waited:.1f}"
          "   (1 guard, 2 candidates, 3 delta staged, 4 NMS, 5 arg-max map staged, 6 walk, 7 slots assigned; inside the NMS: 11 ranked, 12 sorted, 13 diagonal blocks; wavefront: 14 word 0 ready to resolve, 15 published, 16 word 1 ready, 17 published, 18 word 2 published, 19 last word published)")
    if k3:
        print("   arg-max kernel, us after its start (first CTA past each point): " + "  ".join(f"p{k} {v:.1f}" for k, v in k3.items())
              + "   (22 rows streamed, 23 CTA reduced, 24 cluster barrier, 25 rank 0 gathered and stored; 121 / 122: the LAST CTA started / had its rows streamed)")
    print(f"   back to back on the stream          {stream:6.1f} us per call   ({img_bytes / stream / 1e3:.0f} GB/s)")
    del bufs, parser
    torch.cuda.empty_cache()
