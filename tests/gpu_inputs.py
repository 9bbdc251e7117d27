"""Synthetic head tensors generated ON THE DEVICE: ``torch.Generator(device).manual_seed(seed)``, the
identical tensor copied to the host for the oracle.

  U  torch.rand                                     R  torch.sigmoid(torch.randn)   ("random-init-like")
  D  U with resp, conf in [0.4, 1) and w, h * 0.08  S  U with resp ** 8 and w, h * 0.3

The one thing the reference leaves undefined is the visiting order of exactly equal root scores (datatest.py:139), so
parity inputs must not contain any: `distinct_root_scores` nudges one of every pair of equal scores above the
threshold by an ulp of `resp` until none is left (and says how many it moved) — never a silent skip.
"""
import numpy as np
import torch


def device_head(g, dist: str, seed: int, B: int, device="cuda", dtype=torch.float32, distinct: bool = True):
    """-> (device tensor [B, C, H, W] of `dtype`, its host copy widened to fp32 — what the oracle parses)."""
    gen = torch.Generator(device=device).manual_seed(int(seed))
    K = g.K
    shape = (B, g.C, g.H, g.W)
    if dist == "R":
        t = torch.sigmoid(torch.randn(shape, device=device, generator=gen))
    else:
        t = torch.rand(shape, device=device, generator=gen)
        if dist == "D":
            t[:, :2 * K] = 0.4 + 0.6 * t[:, :2 * K]
            t[:, 4 * K:6 * K] *= 0.08
        elif dist == "S":
            t[:, :K] = t[:, :K] ** 8
            t[:, 4 * K:6 * K] *= 0.3
        elif dist != "U":
            raise ValueError(dist)
    t = t.to(dtype).contiguous()
    host = t.float().cpu().numpy()
    if distinct:                                   # (16-bit heads are full of equal scores: those tests pin the build's own tie rule)
        if distinct_root_scores(host, g):
            assert dtype == torch.float32
            t = torch.from_numpy(host).to(device)
    return t, host


def distinct_root_scores(host: np.ndarray, g, thr: float = None) -> int:
    """Make the root scores delta[0] = resp[0] * conf[0] above the detection threshold pairwise distinct, in place,
    by moving resp[0] of the later cell of an equal pair to the next float; returns the number of cells moved and
    asserts that none is left."""
    thr = np.float32(g.det_thresh if thr is None else thr)
    K, moved = g.K, 0
    for img in host:
        resp, conf = img[0].reshape(-1), img[K].reshape(-1)
        for _ in range(64):
            d = resp * conf
            live = np.flatnonzero(d > thr)
            vals, inv, cnt = np.unique(d[live], return_inverse=True, return_counts=True)
            if (cnt == 1).all():
                break
            for v in np.flatnonzero(cnt > 1):
                for cell in live[inv == v][1:]:
                    resp[cell] = np.nextafter(resp[cell], np.float32(2.0))
                    moved += 1
        else:
            raise AssertionError("could not make the root scores distinct")
        d = resp * conf
        live = d[d > thr]
        assert np.unique(live).size == live.size
    return moved
