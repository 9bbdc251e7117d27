#!/usr/bin/env python
"""Benchmark of the PPN output-parsing path (decode + NMS + limb arg-max + tree parse).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl ours|reference]

One step = one pass of the whole path over one batch of synthetic head tensors per GPU
(BASELINE.json configs[1]: MPII 16-part, 384x384, 12x12 grid, 9x9 window, batch 512).  N > 1 is
launched by torchrun, one rank per GPU; every rank parses its own batch (weak scaling, images
are independent) and the packed poses are all-gathered over NCCL inside the timed region.
Rank 0 prints ONE JSON line (see README / DESIGN.md for the keys).

`--impl reference` times the reference's CPU algorithm (the numpy port in oracle/, which follows
datatest.py line by line and is checked against the reference's own outputs) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = {"cfg1": 1, "cfg2": 512, "cfg3": 1024, "cfg4": 256, "native": 64}
DIST = {"cfg1": "U", "cfg2": "U", "cfg3": "D", "cfg4": "U", "native": "U"}
WORKLOAD = {
    "cfg1": "MPII 16-part PPN, 384x384 (12x12 grid, 9x9 limb window), batch 1 (latency)",
    "cfg2": "MPII 16-part PPN, 384x384 (12x12 grid, 9x9 limb window), batch 512 synthetic head tensors per GPU",
    "cfg3": "COCO 18-part PPN, 512x512 (16x16 grid, 9x9 window), batch 1024, dense crowd",
    "cfg4": "18-part PPN, 768x768 (24x24 grid, 11x11 window), batch 256",
    "native": "reference-native 18-part PPN, 384x384 (24x24 grid, 21x21 window), batch 64",
}
METRIC = "PPN decode+NMS+parse images/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ----------------------------------------------------------------------------------------------
# clocks: sampled DURING the run with NVML (falls back to nvidia-smi)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, cuda_index: int, period_s: float = 0.01):
        self.period = period_s
        self.samples = []          # (t, sm_mhz, reasons_bits, power_w)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self.h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + uuid)
                except Exception:
                    self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                                     # pragma: no cover
            self.err = repr(e)
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((time.perf_counter(), float(mhz), int(bits), pw))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.h is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)

    def summary(self, t0=None, t1=None):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + getattr(self, "err", "")}
        sel = [s for s in self.samples if (t0 is None or s[0] >= t0) and (t1 is None or s[0] <= t1)]
        window = "timed region"
        if len(sel) < 3:                       # a short timed region: use every sample taken under load
            sel, window = self.samples, "warm-up + timed region"
        reasons = set()
        for _, _, bits, _ in sel:
            for bit, name in self.REASONS.items():
                if bits & bit:
                    reasons.add(name)
        mhz = [s[1] for s in sel]
        pw = [s[3] for s in sel if s[3] is not None]
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(sel), "window": window,
                "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------------------------
# reference arm: the CPU algorithm on all host cores
# ----------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return 0
    from oracle import cpu_bench, ppn_oracle as O
    from pytorch_pose_proposal_network_b200.config import PRESETS
    cfg = PRESETS[args.config]()
    g = O.Geometry.of(cfg)
    cores = cpu_bench.host_cores()
    per_step = max(cores * 4, 16)                    # images parsed per step, split over the cores
    pool = cpu_bench.CpuPool(g, dist=DIST[args.config], seed=0, n_images=per_step, cores=cores)
    try:
        for _ in range(max(args.warmup, 1)):
            pool.one_pass()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.one_pass()
        dt = time.perf_counter() - t0
    finally:
        pool.close()
    value = per_step * args.steps / dt
    sample = (f"{per_step} synthetic images ({DIST[args.config]}, seed 0) per step, numpy port of datatest.py "
              f"(oracle/ppn_oracle.parse_head_like_reference), process pool over {cores} cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[args.config], "preset": args.config, "images_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PackedHumans, PoseParser
    from pytorch_pose_proposal_network_b200.sharded import PoseGatherer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = PRESETS[args.config]()
    B = args.batch or BATCH[args.config]
    K, W = args.steps, max(args.warmup, 3)
    for kv in args.tune:
        k, v = kv.split("=")
        _lib.tune(**{k.replace(".", "_"): int(v)})

    # humans per image kept in the packed output; H*W can never overflow.  For N > 1 the gather
    # ships a trimmed stride (checked against the true counts after the run).
    parser = PoseParser(cfg, device=dev, max_humans=args.max_humans or None)
    per_image = args.gather_entries or 6 * cfg.K   # (human, part) entries shipped per image on average (overflow is checked)
    cap_entries = B * per_image
    gatherer = PoseGatherer(parser, B, cap_entries, group_steps=args.gather_every) if world > 1 else None

    # distinct input batches, rotated so that no step finds its input in the 126 MB L2
    head_dtype = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[args.head_dtype]
    elem = 4 if args.head_dtype == "f32" else 2
    batch_bytes = B * cfg.C * cfg.HW * elem
    n_buf = max(2, min(6, -(-(1 << 30) // batch_bytes)))
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    bufs = []
    for _ in range(n_buf):
        t = torch.rand(B, cfg.C, cfg.H, cfg.W, device=dev, generator=gen)
        if DIST[args.config] == "D":
            t[:, :2 * cfg.K] = 0.4 + 0.6 * t[:, :2 * cfg.K]
            t[:, 4 * cfg.K:6 * cfg.K] *= 0.08
        bufs.append(t.to(head_dtype))               # a 16-bit head: same values rounded once, widened exactly by the kernels
        del t
    outs = [parser.alloc_output(B) for _ in range(2)]

    def step(i):
        # the inputs have been resident in HBM since before the timed region: the parser may overlap
        # consecutive steps (PPN_FLAG_INPUT_COMPLETE); results still complete in step order
        if gatherer is not None:                    # poses go straight into the gather's group buffer (dense records);
            return gatherer.parse(bufs[i % n_buf], out=outs[i % 2], input_complete=not args.no_step_overlap)
        return parser.parse(bufs[i % n_buf], out=outs[i % 2], input_complete=not args.no_step_overlap)

    def drain():
        if gatherer is not None:
            gatherer.finish()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank).start()
    # warm-up: W steps, then keep the GPU under the same load for ~0.4 s so that the clock
    # sampler sees loaded clocks even when the timed region is only milliseconds long
    t_w = time.perf_counter()
    for i in range(W):
        step(i)
    drain()
    torch.cuda.synchronize(dev)
    # the number of extra steps must be the SAME on every rank (each submits collectives): rank 0
    # sizes it from its own warm-up time and broadcasts it
    per_step = max((time.perf_counter() - t_w) / W, 1e-5)
    n_extra = torch.tensor([min(20000, int(args.settle_s / per_step) + 1)], device=dev, dtype=torch.int64)
    if world > 1:
        dist.broadcast(n_extra, src=0)
    extra = int(n_extra.item())
    for i in range(extra):
        step(i)
    drain()
    torch.cuda.synchronize(dev)

    # optional: replay the steps from a CUDA graph (one graph = `group` consecutive steps, so that the
    # rotation of input and output buffers is part of it); the remainder of K runs eagerly
    graph, group = None, 0
    if args.cuda_graph and world == 1:
        group = 2 * n_buf                           # a multiple of both rotations
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for i in range(group):
                    step(i)
        torch.cuda.current_stream(dev).wait_stream(side)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize(dev)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t0 = time.perf_counter()
    ev0.record()
    last = None
    if graph is not None:
        for _ in range(K // group):
            graph.replay()
        for i in range(K - K % group, K):
            last = step(i)
        last = outs[(K - 1) % 2] if last is None else last
    else:
        for i in range(K):
            last = step(i)
    drain()                                         # the timed region ends when every gather has landed
    ev1.record()
    host_issue_ms = (time.perf_counter() - t0) * 1e3 / K     # host time to ENQUEUE a step (GPU runs behind)
    sync_all()
    t1 = time.perf_counter()
    elapsed_ms = ev0.elapsed_time(ev1)

    # Spread of the measurement (SURVEY §8d asks for median and min): the same K-step region four more times.
    # `value` stays the FIRST region's (exactly K steps after the warm-up); these are reported beside it.
    repeat_ms = [elapsed_ms / K]
    if graph is None:
        for _ in range(4):
            ra, rb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sync_all()
            ra.record()
            for i in range(K):
                step(i)
            drain()
            rb.record()
            sync_all()
            r_ms = ra.elapsed_time(rb)
            if world > 1:
                tm = torch.tensor([r_ms], device=dev, dtype=torch.float64)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                r_ms = float(tm.item())
            repeat_ms.append(r_ms / K)
    t1 = time.perf_counter()

    # Per-kernel durations: the same K steps again, same inputs, with the library recording CUDA
    # events around every kernel on its stream (ppn_profile_*).  Bracketing a kernel with events
    # forbids the overlapped launch chain the timed region above uses, so this pass runs the three
    # kernels back to back; its step time is reported too and is NOT the headline value.
    _lib.profile_enable(True)
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Kp = min(K, 4096)
    sync_all()
    evp0.record()
    for i in range(Kp):
        last = step(i)
    drain()
    evp1.record()
    sync_all()
    t1 = time.perf_counter()
    profiled_ms_per_step = evp0.elapsed_time(evp1) / Kp
    stage_ms, n_prof = _lib.profile_read()
    _lib.profile_enable(False)
    clocks = sampler.summary(t0, t1)

    if world > 1:
        tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
        repeat_ms[0] = elapsed_ms / K
    value = world * B * K / (elapsed_ms * 1e-3)

    # sanity on what was produced (outside the timed region): counts must fit the gathered stride
    counts = last.count.cpu()
    humans_per_image = float(counts.float().mean())
    if gatherer is not None:
        for r in range(world):
            rec = gatherer.records_of(r)
            if rec["overflow"]:
                raise SystemExit(f"bench.py: rank {r} produced {rec['total']} pose entries, more than the {cap_entries} "
                                 f"shipped per step; raise --gather-entries")
        mine = gatherer.records_of(rank)            # what every rank received from this rank == what it produced
        assert int(mine["count"].sum()) == int(counts.sum()), "gathered humans differ from the local result"

    # ---- end to end through the public host-buffer call: H2D + kernels + D2H every step ----
    host_in = torch.empty(B, cfg.C, cfg.H, cfg.W, dtype=head_dtype, pin_memory=True)
    host_in.copy_(bufs[0])
    host_out = parser.alloc_output(B, device="cpu", pin=True)
    Ke = max(1, min(K, args.e2e_steps))
    for _ in range(2):
        parser.parse_host(host_in, out=host_out)
    sync_all()
    te0 = time.perf_counter()
    for _ in range(Ke):
        parser.parse_host(host_in, out=host_out)      # synchronous: returns with results on the host
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - te0
    if world > 1:
        tmax = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    d2h = sum(getattr(host_out, f).numel() * getattr(host_out, f).element_size()
              for f in ("count", "root_cell", "part_cell", "part_score", "part_box"))
    e2e = {"value": world * B * Ke / e2e_s, "unit": "images/s", "h2d_bytes_per_step": batch_bytes,
           "d2h_bytes_per_step": d2h, "steps": Ke, "api": "PoseParser.parse_host -> ppn_parse_host (pinned host buffers)"}
    sampler.stop()

    # ---- roofline of the dominant kernel (limb arg-max), timed inside the timed region ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    limb_bytes = B * cfg.E * cfg.S * cfg.HW * elem + B * cfg.E * cfg.HW * 2    # read once + uint16 map written
    traffic = None                                   # DRAM bytes per launch from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(args.config)
        if rec and rec.get("images") == B and args.head_dtype == "f32":
            traffic = rec["dram_bytes_read"] + rec["dram_bytes_write"]
    k3_ms = stage_ms["limb_argmax"] / max(n_prof, 1)
    achieved = limb_bytes / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "limb_argmax", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": limb_bytes, "avg_launch_ms": k3_ms,
                "timing": f"CUDA events around each kernel on its stream, {n_prof} steps run right after the timed region "
                          "(events forbid the overlapped launch chain, so kernels run back to back in this pass)",
                "stage_ms_per_step": ({"limb_argmax": stage_ms["limb_argmax"] / max(n_prof, 1),
                                       "parse_fused": stage_ms["tree_parse"] / max(n_prof, 1)}
                                      if parser.launches_per_parse(B) == 2 else
                                      {"limb_argmax": stage_ms["limb_argmax"] / max(n_prof, 1),
                                       "decode_nms": stage_ms["nms"] / max(n_prof, 1),
                                       "tree_parse": stage_ms["tree_parse"] / max(n_prof, 1)}),
                "serial_ms_per_step": profiled_ms_per_step,
                "pipeline_gbs": batch_bytes / (elapsed_ms / K * 1e-3) / 1e9,
                "pipeline_frac": batch_bytes / (elapsed_ms / K * 1e-3) / 1e9 / peak}

    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[args.config], "preset": args.config, "images_per_gpu_per_step": B,
                   "K": cfg.K, "E": cfg.E, "grid": [cfg.H, cfg.W], "window": [cfg.sH, cfg.sW],
                   "bytes_per_image": cfg.C * cfg.HW * elem, "head_dtype": args.head_dtype,
                   "input_distribution": DIST[args.config],
                   "l2": f"{n_buf} distinct input batches of {batch_bytes / 1e6:.0f} MB rotated (each larger than L2)",
                   "humans_per_image": humans_per_image, "extra_warmup_steps": extra,
                   "host_issue_ms_per_step": host_issue_ms, "cuda_graph": bool(graph is not None),
                   "step_overlap": "off" if args.no_step_overlap else
                   "PPN_FLAG_INPUT_COMPLETE: inputs resident before the timed region, so step i+1's arg-max may start while "
                   "step i's tree parse finishes; steps complete in order",
                   "pose_gather": "none (1 GPU)" if world == 1 else
                   f"every step: the parse kernel writes dense (human, part) entries itself (cap {per_image}/image avg); every "
                   f"{args.gather_every} steps one async NCCL all_gather of {args.gather_every * gatherer.nbytes / 1e6:.2f} MB "
                   f"per rank, overlapped with the following steps; all gathers complete inside the timed region"},
        "roofline": roofline, "e2e": e2e, "clocks": clocks,
        "gpu_launches": parser.launches_per_parse(B) * K,
        "repeats": {"ms_per_step": repeat_ms, "median_ms_per_step": sorted(repeat_ms)[len(repeat_ms) // 2],
                    "min_ms_per_step": min(repeat_ms), "note": "the timed K-step region run five times; `value` is the first"},
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline(cfg, args):
    """The reference-shaped numpy port on all host cores, bounded to ~args.cpu_budget_s seconds."""
    from oracle import cpu_bench, ppn_oracle as O
    g = O.Geometry.of(cfg)
    cores = cpu_bench.host_cores()
    dist = DIST[args.config]
    one_core_ips, one_core_ms, n1 = cpu_bench.time_single_core(g, dist, seed=0, n_images=16, budget_s=min(4.0, args.cpu_budget_s / 3))
    per_pass = max(4 * cores, 16)
    pool = cpu_bench.CpuPool(g, dist=dist, seed=0, n_images=per_pass, cores=cores)
    try:
        pool.one_pass()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < args.cpu_budget_s * 0.6:
            pool.one_pass()
            n += per_pass
        dt = time.perf_counter() - t0
    finally:
        pool.close()
    c_ips = cpu_bench.time_c_port(g, dist, seed=0, n_images=max(2 * cores, 16), threads=cores, budget_s=min(3.0, args.cpu_budget_s / 4))
    return {"value": n / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{n} images ({per_pass}-image synthetic sample, dist {dist}, repeated for {dt:.1f} s) through the numpy port "
                      f"of datatest.get_humans_by_feature, one process per core",
            "one_core_images_per_s": one_core_ips, "one_core_ms_per_image": one_core_ms,
            "c_port_images_per_s": c_ips, "c_port_threads": cores}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(BATCH))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the config's)")
    ap.add_argument("--max-humans", type=int, default=0, help="slots per image in the packed output (default H*W)")
    ap.add_argument("--gather-entries", type=int, default=0,
                    help="average (human, part) entries per image shipped by the N>1 gather (default 6*K; overflow is detected)")
    ap.add_argument("--gather-every", type=int, default=32, help="steps per pose all_gather (N > 1)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--settle-s", type=float, default=0.4)
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--head-dtype", default="f32", choices=["f32", "f16", "bf16"],
                    help="element type of the head tensor (the reference's is f32; the arithmetic is fp32 either way)")
    ap.add_argument("--cuda-graph", action="store_true", help="replay the timed steps from a CUDA graph (1 GPU)")
    ap.add_argument("--no-step-overlap", action="store_true",
                    help="do not pass PPN_FLAG_INPUT_COMPLETE (each step's kernels wait for the previous step's)")
    ap.add_argument("--tune", action="append", default=[], help="library knob, e.g. argmax.stages=6")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
