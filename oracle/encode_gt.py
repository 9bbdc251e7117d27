"""Ground-truth encoder: people -> target grids.  TEST INFRASTRUCTURE (KAT generation).

The inverse of the parser, restated from the reference's ``KeypointsDataset.__getitem__``
(/root/reference/dataset.py:98-152): a labelled part at pixel (px, py) switches on the cell
that contains it, stores the offset inside the cell and its box size as fractions, and each
limb whose two ends are labelled sets a one-hot in the source cell's displacement window.
The reference's only self-check of its parser is parsing these targets back
(datatest.py:403-412); ``tests`` do the same through the CUDA path.
"""
from __future__ import annotations

import numpy as np


def encode_people(people, g, edges, part_size=None):
    """people: list of dicts {'box': (cx, cy, w, h), 'points': {part_id: (px, py)}} in pixels.

    Returns a head tensor [C,H,W] fp32 with resp = delta targets, conf = 1, x, y, w, h targets
    and the one-hot limb block, i.e. exactly what ``datatest.py:405-412`` feeds the parser.
    """
    K, E, H, W, sH, sW = g.K, g.E, g.H, g.W, g.sH, g.sW
    gridW, gridH = int(g.inW / g.W), int(g.inH / g.H)
    delta = np.zeros((K, H, W), np.float32)
    tx, ty, tw, th = (np.zeros((K, H, W), np.float32) for _ in range(4))
    te = np.zeros((E, sH, sW, H, W), np.float32)
    for person in people:
        cx, cy, bw, bh = person["box"]
        psize = part_size if part_size is not None else max(bw, bh) / 8.0
        pts = dict(person["points"])
        if bw > 0 and bh > 0:
            pts[0] = (cx, cy)                                  # dataset.py:111-117: part 0 is the box centre
        for k, (px, py) in pts.items():
            fx, fy = px / gridW, py / gridH
            ix, iy = int(fx), int(fy)
            if 0 <= iy < H and 0 <= ix < W:                    # dataset.py:130-135
                delta[k, iy, ix] = 1
                tx[k, iy, ix] = fx - ix
                ty[k, iy, ix] = fy - iy
                tw[k, iy, ix] = (bw if k == 0 else psize) / g.inW
                th[k, iy, ix] = (bh if k == 0 else psize) / g.inH
        for ei, (s, t) in enumerate(edges):                    # dataset.py:137-152
            if s not in pts or t not in pts:
                continue
            sy, sx = int(pts[s][1] / gridH), int(pts[s][0] / gridW)
            dy = int(pts[t][1] / gridH) - sy + sH // 2
            dx = int(pts[t][0] / gridW) - sx + sW // 2
            if not (0 <= sy < H and 0 <= sx < W) or not (0 <= dy < sH and 0 <= dx < sW):
                continue
            te[ei, dy, dx, sy, sx] = 1
    conf = np.ones_like(delta)
    return np.concatenate([delta, conf, tx, ty, tw, th, te.reshape(E * sH * sW, H, W)], axis=0).astype(np.float32)
