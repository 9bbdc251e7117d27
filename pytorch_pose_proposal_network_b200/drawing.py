"""Drawing of parsed humans from the packed result — the consumer right after the path.

The reference's ``draw_humans`` (/root/reference/datatest.py:162-232, called by the webcam loop at
rt_test.py:138-145) walks per-human dicts in Python and computes every rectangle, keypoint and limb segment with
numpy scalars.  Here those primitives come from one kernel launch over the packed result
(``ppn_skeleton`` / :meth:`PoseParser.skeleton`) and this module only hands them to PIL, in the reference's order
and with its colours, so the image is pixel-identical to the reference's.
"""
from __future__ import annotations

import numpy as np

from .config import COLOR_MAP, DIRECTED_GRAPHS


def insertion_order(directed_graphs=DIRECTED_GRAPHS):
    """Part ids in the order the reference inserts them into a human's dict (datatest.py:104-125): the root, then
    the targets of the track orders, first insertion wins.  A human's own key order is this list without its
    absent parts (presence is prefix-closed along every track order)."""
    order = [0]
    for _, ts in directed_graphs:
        for t in ts:
            if t not in order:
                order.append(t)
    return order


def draw_skeletons(pil_image, rect, keypoint, segment, keypoint_names, edges, visbbox=False, part_box=None,
                   directed_graphs=DIRECTED_GRAPHS, color_map=COLOR_MAP):
    """Draw ONE image's humans.  rect [n, 4] int, keypoint [n, K, 2] (x, y), segment [n, E, 4] (bx, by, ex, ey) as
    :meth:`PoseParser.skeleton` returns them for the image's first ``count`` slots (NaN = absent).  ``visbbox``
    draws the parts' boxes instead of dots and needs ``part_box`` [n, K, 4] (ymin, xmin, ymax, xmax)."""
    from PIL import ImageDraw
    drawer = ImageDraw.Draw(pil_image)
    order = insertion_order(directed_graphs)
    rect, keypoint, segment = np.asarray(rect), np.asarray(keypoint), np.asarray(segment)
    colours = [color_map[n] for n in keypoint_names]
    r = 2
    for i in range(rect.shape[0]):
        for k in order:
            x, y = keypoint[i, k]
            if x != x:                                     # absent part
                continue
            if k == 0:                                     # the instance: a two-pixel rectangle (datatest.py:186-191)
                xmin, ymin, xmax, ymax = (int(v) for v in rect[i])
                drawer.rectangle(xy=[xmin, ymin, xmax, ymax], fill=None, outline=colours[0])
                if xmax - xmin >= 2 and ymax - ymin >= 2:  # (the reference raises inside PIL on a box this small)
                    drawer.rectangle(xy=[xmin + 1, ymin + 1, xmax - 1, ymax - 1], fill=None, outline=colours[0])
            elif visbbox:
                ymin, xmin, ymax, xmax = part_box[i, k]
                drawer.rectangle(xy=[xmin, ymin, xmax, ymax], fill=None, outline=colours[k])
            else:
                drawer.ellipse((x - r, y - r, x + r, y + r), fill=colours[k])
        for e, (s, _t) in enumerate(edges):
            bx, by, ex, ey = segment[i, e]
            if bx == bx:
                drawer.line([bx, by, ex, ey], fill=colours[s], width=2)
    return pil_image
