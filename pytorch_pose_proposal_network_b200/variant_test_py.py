"""The parser as the reference's older ``test.py`` calls it (/root/reference/test.py:147-252).

``test.py`` carries a second, stale copy of the parser with other constants — detection threshold
0.09 (test.py:159), NMS threshold 0.5 (test.py:184), ``min_num_keypoints=-1`` (root-only humans are
kept), ``resp * conf`` done inside (test.py:162), geometry taken from ``model`` instead of module
globals, the half-window subtracted the other way round (test.py:211-212), and only ``humans`` is
returned.  Here these are parameters of the same kernels, not a second code path.
"""
from __future__ import annotations

import numpy as np
import torch

from .config import DIRECTED_GRAPHS, PPNConfig
from .datatest import _dev32, _parser_for


def get_humans_by_feature(model, resp, conf, x, y, w, h, e, detection_thresh=0.09, min_num_keypoints=-1):
    """Same arguments as test.py:159; ``model`` supplies ``insize``, ``outsize``, ``local_grid_size``."""
    resp_d, conf_d, e_d = _dev32(resp), _dev32(conf), _dev32(e)
    K, H, W = resp_d.shape
    if (W, H) != tuple(model.outsize):
        raise ValueError(f"arrays are {W}x{H} cells but model.outsize is {tuple(model.outsize)}")
    cfg = PPNConfig(K=K, E=e_d.shape[0], insize=tuple(model.insize), outsize=(W, H),
                    local_grid_size=tuple(model.local_grid_size), directed_graphs=DIRECTED_GRAPHS,
                    detection_thresh=detection_thresh, nms_thresh=0.5, min_num_keypoints=min_num_keypoints,
                    swap_window_offsets=True)
    head = torch.cat([resp_d, conf_d, _dev32(x), _dev32(y), _dev32(w), _dev32(h), e_d.reshape(-1, H, W)], dim=0).unsqueeze(0)
    humans, _ = _parser_for(cfg).parse(head).humans(0)
    return humans
