"""Prediction records for the reference's AP evaluation, built from the packed GPU result.

``datatest.evaluation`` (/root/reference/datatest.py:278-369) first turns every image's
``(humans, scores)`` into an MPII-style frame ``{"image": [fname], "annorect": [...]}``
(datatest.py:298-328) — one dict per person with the root box, the root score and, for each of the
K-1 joints, the centre of its box and its score (zeros when absent) — and then hands the frames to
the poseval fork.  :func:`pred_frame` builds exactly that frame from :class:`..parser.PackedHumans`
and the centres computed on the GPU (``PoseParser.part_centres`` -> ``ppn_part_centres``), without
going through per-human dicts.  The matching itself (``evaluateAP``) stays the reference's.
"""
from __future__ import annotations

import numpy as np


def pred_frame(fname, packed_np, centres_np, b: int):
    """packed_np: ``PackedHumans.numpy()``; centres_np: ``part_centres(...).cpu().numpy()``.

    Value types follow the reference: numpy fp32 scalars for present parts, Python ints 0 for
    absent ones, joint ids ``num - 1``."""
    n = int(packed_np["count"][b])
    cell, score, box = packed_np["part_cell"][b], packed_np["part_score"][b], packed_np["part_box"][b]
    K = cell.shape[1]
    rects = []
    for i in range(n):
        y1, x1, y2, x2 = box[i, 0]
        pp = {"x1": [x1], "y1": [y1], "x2": [x2], "y2": [y2], "score": [score[i, 0]], "annopoints": [{"point": []}]}
        points = pp["annopoints"][0]["point"]
        for num in range(1, K):
            if cell[i, num] >= 0:
                y, x, s = centres_np[b, i, num, 0], centres_np[b, i, num, 1], score[i, num]
            else:
                y, x, s = 0, 0, 0
            points.append({"id": [num - 1], "x": [x], "y": [y], "score": [s]})
        rects.append(pp)
    return {"image": [fname], "annorect": rects}


def pred_frames(fnames, packed, centres):
    """One frame per image of the batch, ready for ``eval_helpers.load_data`` (datatest.py:350)."""
    packed_np = packed.numpy()
    centres_np = centres.cpu().numpy() if hasattr(centres, "cpu") else np.asarray(centres)
    return [pred_frame(f, packed_np, centres_np, b) for b, f in enumerate(fnames)]
