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


def encode_targets(keypoints, bbox, visible, size, K, edges, insize, outsize, window):
    """The reference's training-target encoder, restated (dataset.py:98-185).  TEST INFRASTRUCTURE.

    keypoints fp32 [n, K-1, 2] (x, y); bbox float64 [n, 4] (cx, cy, w, h); visible bool [n, K-1];
    size float64 [n] (side of a part's box) — the types the reference's transform pipeline delivers
    (aug.py:138-160: ``torch.from_numpy(keypoints)``, ``torch.from_numpy(np.asarray(bbox))``, JSON floats).
    Returns [delta, weight, weight_ij, tx, ty, tx_half, ty_half, tw, th, te] like ``__getitem__`` does
    after the image (dataset.py:198).  Arithmetic as the reference's: cell coordinates by fp32 division
    (torch fp32 tensor / int, dataset.py:125-126), offsets by fp32 subtraction, sizes by float64 division
    rounded once on the store into the fp32 grid (dataset.py:134-135); later people overwrite earlier ones.
    The reference's window slicing (dataset.py:163-167) only works for odd square windows; so does this.
    """
    inW, inH = insize
    outW, outH = outsize
    sW, sH = window
    if sW != sH or sW % 2 == 0:
        raise ValueError("the reference's encoder needs an odd square limb window")
    gridW, gridH = int(inW / outW), int(inH / outH)
    E = len(edges)
    f32 = np.float32
    delta = np.zeros((K, outH, outW), f32)
    tx, ty, tw, th = (np.zeros((K, outH, outW), f32) for _ in range(4))
    te = np.zeros((E, sH, sW, outH, outW), f32)
    keypoints = np.asarray(keypoints, f32).reshape(-1, K - 1, 2)
    bbox = np.asarray(bbox, np.float64).reshape(-1, 4)
    for p in range(bbox.shape[0]):
        cx, cy, w, h = bbox[p]
        pts = np.concatenate([np.array([[f32(cx), f32(cy)]], f32), keypoints[p]], axis=0)        # dataset.py:112
        labeled = np.concatenate([[bool(w > 0 and h > 0)], np.asarray(visible[p], bool)])          # dataset.py:115-118
        cells = np.full((K, 2), -(1 << 30), np.int64)
        for k in range(K):
            if not labeled[k]:
                continue
            fx, fy = f32(pts[k, 0] / f32(gridW)), f32(pts[k, 1] / f32(gridH))                      # fp32 division
            ix, iy = int(fx), int(fy)                                                              # truncation
            cells[k] = (iy, ix)
            if 0 <= iy < outH and 0 <= ix < outW:
                delta[k, iy, ix] = 1
                tx[k, iy, ix] = f32(fx - f32(ix))
                ty[k, iy, ix] = f32(fy - f32(iy))
                tw[k, iy, ix] = f32((w if k == 0 else float(size[p])) / inW)                       # float64 division
                th[k, iy, ix] = f32((h if k == 0 else float(size[p])) / inH)
        for ei, (s, t) in enumerate(edges):                                                        # dataset.py:137-152
            if not (labeled[s] and labeled[t]):
                continue
            iy, ix = cells[s]
            jy, jx = cells[t][0] - iy + sH // 2, cells[t][1] - ix + sW // 2
            if iy < 0 or ix < 0 or iy >= outH or ix >= outW:
                continue
            if jy < 0 or jx < 0 or jy >= sH or jx >= sW:
                continue
            te[ei, jy, jx, iy, ix] = 1
    # max(delta_s at the cell, delta_t at the displaced cell), dataset.py:154-170
    mx = np.zeros((E, sH, sW, outH, outW), f32)
    o = sH // 2
    for ei, (s, t) in enumerate(edges):
        pad = np.pad(delta[t], o)
        for dy in range(sH):
            for dx in range(sW):
                mx[ei, dy, dx] = pad[dy:dy + outH, dx:dx + outW]
        mx[ei][:, :, delta[s] != 0] = 1.0
    small = f32(0.0005)
    weight_ij = np.minimum(mx + np.where(mx < 0.5, small, f32(0)), f32(1.0)).astype(f32)           # dataset.py:173-175
    weight = np.minimum(delta + np.where(delta < 0.5, small, f32(0)), f32(1.0)).astype(f32)        # dataset.py:178-180
    half = np.where(delta < 0.5, f32(0.5), f32(0)).astype(f32)
    return [delta, weight, weight_ij, tx, ty, (tx + half).astype(f32), (ty + half).astype(f32), tw, th, te]


def random_people(rng, n, K, insize, spread=1.15, p_visible=0.75):
    """Synthetic annotation set for one image in the reference's post-transform types: some points fall
    outside the image (also at negative coordinates, where int() truncates towards zero), some boxes are
    empty (part 0 unlabelled), people overlap so that later ones overwrite earlier cells."""
    inW, inH = insize
    lo = (1.0 - spread) / 2.0
    keypoints = ((rng.random((n, K - 1, 2)) * spread + lo) * np.array([inW, inH])).astype(np.float32)
    bbox = np.stack([(rng.random(n) * spread + lo) * inW, (rng.random(n) * spread + lo) * inH,
                     rng.random(n) * inW * 0.5, rng.random(n) * inH * 0.5], axis=1)
    bbox[rng.random(n) < 0.15, 2] = 0.0
    visible = [rng.random(K - 1) < p_visible for _ in range(n)]
    size = (rng.random(n) * 40 + 4).tolist()
    return keypoints, bbox, visible, size
