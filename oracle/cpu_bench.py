"""Time the reference-shaped CPU port on the host cores.  TEST/BENCH INFRASTRUCTURE.

Used only by bench.py's ``cpu_baseline`` leg and by ``bench.py --impl reference``.  What is
timed per image is exactly the reference's host work after the device->host copy
(rt_test.py:122-133): slice the seven groups, ``resp * conf``, ``get_humans_by_feature`` —
through ``oracle/ppn_oracle.parse_head_like_reference`` (single-threaded numpy with per-human
Python loops, like the reference).  All cores are used the way BASELINE.md §4 prescribes: a
process pool over images, one process per core.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_state = {}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _init(geom_kwargs, dist, seed, n_images):
    from oracle import ppn_oracle as O, synth
    g = O.Geometry(**geom_kwargs)
    _state["g"] = g
    _state["O"] = O
    # every worker regenerates the same sample locally (no pickling of MBs per task); a pass over more images than
    # the sample holds cycles through it (bench.py's reference arm: 512 images per step from a 32-image sample)
    _state["head"] = synth.make_head(g, dist, seed, B=n_images)
    _state["n"] = n_images


def _work(span):
    O, g, head = _state["O"], _state["g"], _state["head"]
    lo, hi = span
    n_h = 0
    n = _state["n"]
    for b in range(lo, hi):
        humans, _ = O.parse_head_like_reference(head[b % n], g)
        n_h += len(humans)
    return n_h


def geometry_kwargs(g):
    return dict(K=g.K, E=g.E, inW=g.inW, inH=g.inH, W=g.W, H=g.H, sW=g.sW, sH=g.sH, graphs=g.graphs,
                det_thresh=g.det_thresh, nms_thresh=g.nms_thresh, min_kp=g.min_kp, off_h=g.off_h, off_w=g.off_w)


class CpuPool:
    """A pool of `cores` processes holding the same `n_images`-image sample; each pass parses the
    whole sample once, split evenly across the processes."""

    def __init__(self, g, dist="U", seed=0, n_images=None, cores=None, sample_images=None):
        self.cores = cores or host_cores()
        self.n_images = n_images or 4 * self.cores
        self.sample_images = min(self.n_images, sample_images or self.n_images)
        self.g = g
        ctx = mp.get_context("spawn")          # the parent may hold a CUDA context: never fork it
        self.pool = ctx.Pool(self.cores, initializer=_init,
                             initargs=(geometry_kwargs(g), dist, seed, self.sample_images))
        per = -(-self.n_images // self.cores)
        self.spans = [(lo, min(lo + per, self.n_images)) for lo in range(0, self.n_images, per)]
        self.pool.map(_work, [(0, 1)] * self.cores)   # start every worker, build its sample

    def one_pass(self) -> int:
        return sum(self.pool.map(_work, self.spans, chunksize=1))

    def close(self):
        self.pool.close()
        self.pool.join()


def time_single_core(g, dist="U", seed=0, n_images=32, budget_s=5.0):
    """images/s of the port on ONE core, median per-image time over up to `budget_s`."""
    from oracle import ppn_oracle as O, synth
    head = synth.make_head(g, dist, seed, B=n_images)
    times = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end and len(times) < 20 * n_images:
        for b in range(n_images):
            t0 = time.perf_counter()
            O.parse_head_like_reference(head[b], g)
            times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return 1.0 / med, med * 1e3, len(times)


def time_c_port(g, dist="U", seed=0, n_images=64, threads=None, budget_s=5.0):
    """images/s of the C restatement with `threads` pthreads (dense arg-max, an extra data point)."""
    from oracle import c_oracle, synth
    threads = threads or host_cores()
    head = synth.make_head(g, dist, seed, B=n_images)
    c_oracle.parse_batch(head[:2], g, 1)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        c_oracle.parse_batch(head, g, threads)
        n += n_images
    return n / (time.perf_counter() - t0)
