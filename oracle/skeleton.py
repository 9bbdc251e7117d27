"""numpy restatement of what the reference's draw_humans computes before it draws.  TEST INFRASTRUCTURE.

/root/reference/datatest.py:162-232: per human, the root box truncated to ints (:177-181), the centre of every
present part's box (:200-202) and, for every limb whose two parts are present, the segment between their centres
(:213-221).  Same layout as ppn_skeleton (include/ppn_decode.h).
"""
from __future__ import annotations

import numpy as np


def primitives(part_cell: np.ndarray, part_box: np.ndarray, edges):
    """part_cell [n, K] (-1 absent), part_box [n, K, 4] fp32 (ymin, xmin, ymax, xmax) of ONE image's humans ->
    rect [n, 4] int32 (xmin, ymin, xmax, ymax), keypoint [n, K, 2] fp32 (x, y), segment [n, E, 4] fp32; NaN = absent."""
    n, K = part_cell.shape
    E = len(edges)
    box = np.asarray(part_box, np.float32)
    present = part_cell >= 0
    rect = np.zeros((n, 4), np.int32)
    kp = np.full((n, K, 2), np.nan, np.float32)
    seg = np.full((n, E, 4), np.nan, np.float32)
    for i in range(n):
        ymin, xmin, ymax, xmax = box[i, 0]
        rect[i] = (int(xmin), int(ymin), int(xmax), int(ymax))                 # datatest.py:177-181 with t = 1
        for k in range(K):
            if present[i, k]:
                b = box[i, k]
                kp[i, k] = ((b[1] + b[3]) / np.float32(2), (b[0] + b[2]) / np.float32(2))     # :200-202
        for e, (s, t) in enumerate(edges):
            if present[i, s] and present[i, t]:
                seg[i, e] = (kp[i, s, 0], kp[i, s, 1], kp[i, t, 0], kp[i, t, 1])             # :213-221
    return rect, kp, seg
