"""Drop-in for ``inference`` of the reference's webcam script (/root/reference/rt_test.py:87-147).

Same signature — ``inference(image, model, outsize, local_grid_size)`` — and the same
pre-processing (rt_test.py:94-101, including its quirk of applying 0-1-scale mean/std to 0-255
pixels), but everything after ``model(image)`` stays on the GPU: the un-sliced head tensor goes
straight into :class:`..parser.PoseParser` instead of seven ``.cpu().numpy()`` copies and the
numpy parser (rt_test.py:106-133).

Drawing (``datatest.draw_humans``, rt_test.py:138-145): the rectangles, keypoints and limb segments come from
one more kernel over the packed result (``ppn_skeleton``) and :mod:`.drawing` hands them to PIL in the reference's
order and colours — the returned PIL image is pixel-identical to the reference's.
"""
from __future__ import annotations

import numpy as np
import torch

from .config import DIRECTED_GRAPHS, EDGES, KEYPOINT_NAMES, PPNConfig
from .parser import PoseParser

_parsers = {}


def _parser(model, outsize, local_grid_size, image_size) -> PoseParser:
    names = getattr(model, "keypoint_names", KEYPOINT_NAMES)
    edges = getattr(model, "edges", EDGES)
    insize = tuple(getattr(model, "insize", (image_size, image_size)))
    key = (len(names), len(edges), insize, tuple(outsize), tuple(local_grid_size), torch.cuda.current_device())
    if key not in _parsers:
        cfg = PPNConfig(K=len(names), E=len(edges), insize=insize, outsize=tuple(outsize),
                        local_grid_size=tuple(local_grid_size), directed_graphs=DIRECTED_GRAPHS,
                        detection_thresh=0.15)                                  # rt_test.py:133
        _parsers[key] = PoseParser(cfg)
    return _parsers[key]


def inference(image, model, outsize, local_grid_size, draw=None, image_size=None, return_humans=False):
    """image: PIL image or HWC uint8 array already at the network's input size (rt_test.py:172-189).

    Returns what the reference returns (rt_test.py:135-147): the input frame with the humans drawn on it, a PIL
    image.  The skeletons are drawn from primitives computed on the GPU (``ppn_skeleton``; no per-human dicts are
    built).  ``draw`` — the reference's own ``draw_humans`` — may be passed to do the drawing instead (it is then
    given ``(humans, scores)`` dicts); ``return_humans=True`` returns ``(humans, scores)`` and draws nothing."""
    arr = np.array(image)
    size = image_size or arr.shape[0]
    mean = torch.tensor([0.485, 0.456, 0.406]).cuda().view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).cuda().view(1, 3, 1, 1)
    model.eval()
    x = torch.from_numpy(arr.transpose((2, 0, 1))).cuda().view(1, 3, size, size).float()
    x = x.sub_(mean).div_(std)
    with torch.no_grad():
        output = model(x).detach()
    parser = _parser(model, outsize, local_grid_size, size)
    packed = parser.parse(output.float().contiguous())
    if return_humans:
        return packed.humans(0)
    names, edges = getattr(model, "keypoint_names", KEYPOINT_NAMES), getattr(model, "edges", EDGES)
    from PIL import Image
    raw = x.mul_(std).add_(mean)
    pil = Image.fromarray(np.squeeze(raw.cpu().numpy(), axis=0).astype(np.uint8).transpose(1, 2, 0))
    if draw is not None:
        humans, _ = packed.humans(0)
        return draw(keypoint_names=names, edges=edges, pil_image=pil.copy(), humans=humans, visbbox=False, gridOn=False)
    from .drawing import draw_skeletons
    rect, kp, seg = parser.skeleton(packed, edges=edges)
    n = int(packed.count[0].item())                       # one synchronisation; then only the used slots come home
    n = min(n, packed.R)
    return draw_skeletons(pil.copy(), rect[0, :n].cpu().numpy(), kp[0, :n].cpu().numpy(), seg[0, :n].cpu().numpy(), names, edges)
