"""CPU restatement (numpy) of the reference's pose parser.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
arm may import this module; the product (``pytorch_pose_proposal_network_b200``) never
does and has no CPU fallback.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the reference's own
``datatest.get_humans_by_feature`` / ``non_maximum_suppression`` (imported unmodified from
/root/reference with its missing viz dependencies stubbed, ``oracle/ref_live.py``) on seeded
inputs, checks this restatement against it, and commits the reference's outputs as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks this file against those
vectors wherever the suite runs.

What is restated, with the reference lines each function follows:

=====================  =====================================================
``split_head``          rt_test.py:109-128 (channel groups, limb reshape, squeeze)
``restore_xy``          datatest.py:63-67
``restore_size``        datatest.py:69-71
``boxes``               datatest.py:80-85
``root_candidates``     datatest.py:86-92
``nms``                 datatest.py:134-160
``limb_argmax``         datatest.py:100,113 (the argmax of every window, densely)
``walk``                datatest.py:103-131
``parse_image``         rt_test.py:130-133 + datatest.py:74-132, packed output
``humans_as_dicts``     datatest.py:98-99,129-132 (return types, key order)
``pred_frame``          datatest.py:298-328 (prediction record for the AP evaluation)
=====================  =====================================================

All arithmetic is IEEE fp32 with one rounding per written operation (numpy never fuses),
thresholds are demoted to fp32 as numpy's weak-scalar promotion does.

One thing the reference leaves open: ``score.argsort()[::-1]`` (datatest.py:139) is an
unstable sort, so the order of exactly equal scores is not defined by the reference.  This
restatement (and the CUDA path) orders ties by LARGER candidate index first, which is what
numpy's small-array insertion sort reversed gives; parity inputs avoid duplicate root scores.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32


@dataclass(frozen=True)
class Geometry:
    """Plain-data mirror of the reference's module globals (datatest.py:53-60)."""
    K: int
    E: int
    inW: int
    inH: int
    W: int            # outW
    H: int            # outH
    sW: int
    sH: int
    graphs: tuple     # ((limb ids...), (target part ids...)) per track order, config.py:75-80
    det_thresh: float = 0.15
    nms_thresh: float = 0.3
    min_kp: int = 1
    off_h: Optional[int] = None     # defaults: sW//2 for rows, sH//2 for columns (datatest.py:115-116)
    off_w: Optional[int] = None

    @property
    def gridW(self): return int(self.inW / self.W)
    @property
    def gridH(self): return int(self.inH / self.H)
    @property
    def S(self): return self.sH * self.sW
    @property
    def C(self): return 6 * self.K + self.S * self.E
    @property
    def oh(self): return self.sW // 2 if self.off_h is None else self.off_h
    @property
    def ow(self): return self.sH // 2 if self.off_w is None else self.off_w

    @classmethod
    def of(cls, cfg) -> "Geometry":
        """Build from anything with PPNConfig's attribute names (duck-typed, no import)."""
        return cls(K=cfg.K, E=cfg.E, inW=cfg.inW, inH=cfg.inH, W=cfg.W, H=cfg.H, sW=cfg.sW, sH=cfg.sH,
                   graphs=tuple((tuple(a), tuple(b)) for a, b in cfg.directed_graphs),
                   det_thresh=cfg.detection_thresh, nms_thresh=cfg.nms_thresh,
                   min_kp=cfg.min_num_keypoints, off_h=cfg.off_h, off_w=cfg.off_w)


# --------------------------------------------------------------------------- #
# stage restatements
# --------------------------------------------------------------------------- #
def split_head(out: np.ndarray, g: Geometry):
    """[C,H,W] -> resp, conf, x, y, w, h [K,H,W] and e [E,sH,sW,H,W]  (rt_test.py:109-128)."""
    assert out.dtype == np.float32 and out.shape == (g.C, g.H, g.W), (out.shape, (g.C, g.H, g.W))
    K = g.K
    groups = [out[i * K:(i + 1) * K] for i in range(6)]
    e = out[6 * K:].reshape(g.E, g.sH, g.sW, g.H, g.W)
    return (*groups, e)


def restore_xy(x, y, g: Geometry):
    """datatest.py:63-67 — cell-relative centre -> pixels; two roundings (add, then multiply)."""
    X, Y = np.meshgrid(np.arange(g.W, dtype=f32), np.arange(g.H, dtype=f32))
    return (x + X) * f32(g.gridW), (y + Y) * f32(g.gridH)


def restore_size(w, h, g: Geometry):
    """datatest.py:69-71."""
    return f32(g.inW) * w, f32(g.inH) * h


def boxes(x, y, w, h, g: Geometry) -> np.ndarray:
    """[K,H,W,4] boxes as (ymin, xmin, ymax, xmax)  (datatest.py:80-85)."""
    rx, ry = restore_xy(x, y, g)
    rw, rh = restore_size(w, h, g)
    hw, hh = rw * f32(0.5), rh * f32(0.5)          # '/ 2' is exact, same as '* 0.5'
    return np.stack([ry - hh, rx - hw, ry + hh, rx + hw], axis=-1).astype(f32)


def root_candidates(delta0: np.ndarray, thr) -> np.ndarray:
    """Flat cell ids (h*W + w, ascending) with delta > thr, strict  (datatest.py:89)."""
    return np.flatnonzero(delta0.reshape(-1) > f32(thr)).astype(np.int32)


def iou_one_to_many(b, area_b, others, area_others):
    """datatest.py:145-149 for one box against a set; every op a single fp32 rounding."""
    tl = np.maximum(b[:2], others[:, :2])
    br = np.minimum(b[2:], others[:, 2:])
    d = br - tl
    inter = (d[:, 0] * d[:, 1]) * (tl < br).all(axis=1).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / ((area_b + area_others) - inter)


def nms(bbox: np.ndarray, thresh, score: Optional[np.ndarray] = None, limit: Optional[int] = None) -> np.ndarray:
    """Greedy IoU suppression  (datatest.py:134-160).

    Returns int32 indices into ``bbox`` in visiting order (descending score when ``score``
    is given).  A box is dropped when its IoU with ANY already kept box is >= thresh; NaN
    IoU (0/0 from zero-area boxes) compares false, so such a box is kept.
    """
    bbox = np.asarray(bbox, f32).reshape(-1, 4)
    n = bbox.shape[0]
    if n == 0:
        return np.zeros((0,), np.int32)
    if score is not None:
        score = np.asarray(score, f32)
        # descending score; exact ties -> larger index first (see module docstring)
        order = np.lexsort((-np.arange(n), -score.astype(np.float64)))
    else:
        order = np.arange(n)
    b = bbox[order]
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    thr = f32(thresh)
    kept: List[int] = []
    for i in range(n):
        if kept:
            k = np.asarray(kept)
            if (iou_one_to_many(b[i], area[i], b[k], area[k]) >= thr).any():
                continue
        kept.append(i)
        if limit is not None and len(kept) >= limit:
            break
    return order[np.asarray(kept, np.int64)].astype(np.int32)


def limb_argmax(e: np.ndarray) -> np.ndarray:
    """[E,sH,sW,H,W] -> [E,H,W] index of the FIRST maximum of every (limb, cell) window.

    The reference takes ``np.argmax(e.transpose(0,3,4,1,2)[ei, h, w])`` on demand
    (datatest.py:100,113); this is the same argmax for every cell at once.  numpy's argmax
    returns the first maximum and treats NaN as the maximum.
    """
    E, sH, sW, H, W = e.shape
    return e.reshape(E, sH * sW, H, W).argmax(axis=1).astype(np.int32)


def walk(root_cell: int, delta: np.ndarray, amax: np.ndarray, g: Geometry):
    """Follow every track order from one root cell  (datatest.py:103-127).

    Returns (order, cell): ``order`` lists part ids in first-insertion order (root first),
    ``cell[t]`` is the flat cell of part t or -1.  A chain stops at the first limb whose
    arg-max lands outside the grid or on a cell with delta < thr (strict: equal is accepted).
    """
    thr = f32(g.det_thresh)
    cell = np.full(g.K, -1, np.int32)
    cell[0] = root_cell
    order = [0]
    for eis, ts in g.graphs:
        ih, iw = divmod(int(root_cell), g.W)
        for ei, t in zip(eis, ts):
            a = int(amax[ei, ih, iw])
            jh = ih + a // g.sW - g.oh
            jw = iw + a % g.sW - g.ow
            if jh < 0 or jw < 0 or jh >= g.H or jw >= g.W:
                break
            if delta[t, jh, jw] < thr:
                break
            if t not in order:
                order.append(t)
            cell[t] = jh * g.W + jw
            ih, iw = jh, jw
    return order, cell


@dataclass
class Parsed:
    """Packed result for one image — the layout the CUDA path writes (include/ppn_decode.h)."""
    cand_cell: np.ndarray     # [n_cand] int32, ascending
    keep_idx: np.ndarray      # [n_keep] int32 indices into cand_cell, NMS order
    root_cell: np.ndarray     # [n_h] int32
    part_cell: np.ndarray     # [n_h, K] int32, -1 = absent
    part_score: np.ndarray    # [n_h, K] fp32 (0 where absent)
    part_box: np.ndarray      # [n_h, K, 4] fp32 (0 where absent)
    key_order: list           # per human: part ids in dict insertion order
    amax: np.ndarray          # [E,H,W] int32


def parse_image(out: np.ndarray, g: Geometry, amax: Optional[np.ndarray] = None) -> Parsed:
    """Whole path for one image: rt_test.py:109-133 -> datatest.py:74-132."""
    resp, conf, x, y, w, h, e = split_head(out, g)
    delta = resp * conf                                        # rt_test.py:130
    bbox = boxes(x, y, w, h, g)
    cand = root_candidates(delta[0], g.det_thresh)
    flat_box0 = bbox[0].reshape(-1, 4)
    keep = nms(flat_box0[cand], g.nms_thresh, score=delta[0].reshape(-1)[cand])
    if amax is None:
        amax = limb_argmax(e)
    roots, cells, orders = [], [], []
    for r in cand[keep]:
        order, cell = walk(int(r), delta, amax, g)
        if g.min_kp <= len(order) - 1:                          # datatest.py:129
            roots.append(int(r)); cells.append(cell); orders.append(order)
    n = len(roots)
    part_cell = np.stack(cells).astype(np.int32) if n else np.zeros((0, g.K), np.int32)
    part_score = np.zeros((n, g.K), f32)
    part_box = np.zeros((n, g.K, 4), f32)
    dflat = delta.reshape(g.K, -1)
    bflat = bbox.reshape(g.K, -1, 4)
    for i in range(n):
        for t in np.flatnonzero(part_cell[i] >= 0):
            part_score[i, t] = dflat[t, part_cell[i, t]]
            part_box[i, t] = bflat[t, part_cell[i, t]]
    return Parsed(cand, keep, np.asarray(roots, np.int32), part_cell, part_score, part_box, orders, amax)


def humans_as_dicts(p: Parsed):
    """Packed -> the reference's return value: (humans, scores) lists of dicts (datatest.py:98-132)."""
    humans, scores = [], []
    for i, order in enumerate(p.key_order):
        humans.append({int(t): p.part_box[i, t].copy() for t in order})
        scores.append({int(t): f32(p.part_score[i, t]) for t in order})
    return humans, scores


# --------------------------------------------------------------------------- #
# Reference-shaped entry point: same signature, same per-human Python loops.  This is the
# "port" that bench.py times as the CPU baseline: like the reference it is single-threaded
# numpy with one Python iteration per candidate (NMS) and per limb step (walk), and it takes
# the arg-max of each visited window on demand instead of densely.
# --------------------------------------------------------------------------- #
def get_humans_by_feature(delta, x, y, w, h, e, g: Geometry, detection_thresh=0.15, min_num_keypoints=1):
    """Same arguments and return value as datatest.get_humans_by_feature (datatest.py:74-132)."""
    g = Geometry(**{**g.__dict__, "det_thresh": detection_thresh, "min_kp": min_num_keypoints})
    thr = f32(detection_thresh)
    bbox = boxes(x, y, w, h, g)
    cand_h, cand_w = np.nonzero(delta[0] > thr)
    picked = nms(bbox[0][cand_h, cand_w], g.nms_thresh, score=delta[0][cand_h, cand_w])
    win = e.transpose(0, 3, 4, 1, 2)                     # [E,H,W,sH,sW] view, datatest.py:100
    humans, scores = [], []
    for rh_, rw_ in zip(cand_h[picked], cand_w[picked]):
        human = {0: bbox[0, rh_, rw_]}
        score = {0: delta[0, rh_, rw_]}
        for eis, ts in g.graphs:
            ih, iw = rh_, rw_
            for ei, t in zip(eis, ts):
                dy, dx = np.unravel_index(np.argmax(win[ei, ih, iw]), (g.sH, g.sW))
                jh, jw = ih + dy - g.oh, iw + dx - g.ow
                if not (0 <= jh < g.H and 0 <= jw < g.W) or delta[t, jh, jw] < thr:
                    break
                human[t] = bbox[t, jh, jw]
                score[t] = delta[t, jh, jw]
                ih, iw = jh, jw
        if len(human) - 1 >= min_num_keypoints:
            humans.append(human)
            scores.append(score)
    return humans, scores


def parse_head_like_reference(out: np.ndarray, g: Geometry):
    """rt_test.py:109-133 on one host image: slice, multiply, parse."""
    resp, conf, x, y, w, h, e = split_head(out, g)
    return get_humans_by_feature(resp * conf, x, y, w, h, e, g, g.det_thresh, g.min_kp)


# --------------------------------------------------------------------------- #
# the consumer right after the path: AP-evaluation prediction records
# --------------------------------------------------------------------------- #
def pred_frame(fname, humans, scores, K: int):
    """One image's prediction frame exactly as datatest.evaluation builds it (datatest.py:298-328):
    per person the root box and score, then for joints 1..K-1 the box centre ((a + b) / 2 in fp32)
    and score, or integer zeros when the part is absent."""
    rects = []
    for person, score in zip(humans, scores):
        y1, x1, y2, x2 = person[0]
        pp = {"x1": [x1], "y1": [y1], "x2": [x2], "y2": [y2], "score": [score[0]], "annopoints": [{"point": []}]}
        for num in range(1, K):
            if num in person:
                y = (person[num][0] + person[num][2]) / 2
                x = (person[num][1] + person[num][3]) / 2
                s = score[num]
            else:
                y, x, s = 0, 0, 0
            pp["annopoints"][0]["point"].append({"id": [num - 1], "x": [x], "y": [y], "score": [s]})
        rects.append(pp)
    return {"image": [fname], "annorect": rects}


def canonical(obj):
    """JSON-able form that keeps the value TYPES (numpy fp32 vs Python int), for comparing frames."""
    if isinstance(obj, dict):
        return {k: canonical(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [canonical(v) for v in obj]
    if isinstance(obj, np.floating):
        assert obj.dtype == np.float32, obj.dtype
        return ["f32", int(np.float32(obj).view(np.uint32))]
    if isinstance(obj, (int, np.integer)):
        return ["int", int(obj)]
    if isinstance(obj, str):
        return ["str", obj]
    raise TypeError(type(obj))
