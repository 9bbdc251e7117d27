"""The consumer right after the path: draw_humans of the reference (datatest.py:162-232).

tests/golden/drawings.json holds the sha256 of the images the reference's own draw_humans produced for the humans of
three golden cases (oracle/make_golden.py::drawing_cases).  CPU: the numpy primitives + the package's drawing module
reproduce them.  GPU: the primitives come from ppn_skeleton, and rt_test.inference returns the reference's image.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import skeleton
from tests.golden_util import GOLDEN, load_case

CASES = json.load(open(os.path.join(GOLDEN, "drawings.json")))


def _draw(rect, kp, seg, size, visbbox, part_box):
    from PIL import Image
    from pytorch_pose_proposal_network_b200 import config as pcfg, drawing
    img = drawing.draw_skeletons(Image.new("RGB", (size, size)), rect, kp, seg, pcfg.KEYPOINT_NAMES, pcfg.EDGES,
                                 visbbox=visbbox, part_box=part_box)
    return hashlib.sha256(np.asarray(img).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_primitives_draw_the_references_image(name):
    from pytorch_pose_proposal_network_b200 import config as pcfg
    rec = CASES[name]
    g, out, fx = load_case(name)
    keep = rec["kept"]
    part_cell, part_box = fx["part_cell"][keep], fx["ref_box"][keep]
    rect, kp, seg = skeleton.primitives(part_cell, part_box, pcfg.EDGES)
    assert hashlib.sha256(rect.tobytes() + kp.tobytes() + seg.tobytes()).hexdigest() == rec["primitives_sha256"]
    assert _draw(rect, kp, seg, rec["size"], False, part_box) == rec["sha256"]
    assert _draw(rect, kp, seg, rec["size"], True, part_box) == rec["sha256_visbbox"]


def test_insertion_order_is_the_references_key_order():
    from pytorch_pose_proposal_network_b200 import drawing
    assert drawing.insertion_order() == [0, 15, 13, 1, 3, 5, 2, 4, 6, 17, 14, 7, 9, 11, 8, 10, 12, 16]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_skeleton_primitives(name):
    import torch
    from pytorch_pose_proposal_network_b200 import config as pcfg
    from tests.test_gpu_parity import parser_for
    rec = CASES[name]
    g, out, fx = load_case(name)
    parser = parser_for(g)
    head = torch.from_numpy(np.stack([out, out])).cuda()               # two images: the second checks the batch indexing
    packed = parser.parse(head)
    rect, kp, seg = (t.cpu().numpy() for t in parser.skeleton(packed))
    n = int(packed.count[0])
    assert n == fx["part_cell"].shape[0]
    want = skeleton.primitives(fx["part_cell"], fx["ref_box"], pcfg.EDGES)
    for b in (0, 1):
        for got, ref in zip((rect[b, :n], kp[b, :n], seg[b, :n]), want):
            assert np.array_equal(got.view(np.uint32) if got.dtype == np.float32 else got,
                                  ref.view(np.uint32) if ref.dtype == np.float32 else ref)
        assert (rect[b, n:] == 0).all() and np.isnan(kp[b, n:]).all() and np.isnan(seg[b, n:]).all()
    keep = rec["kept"]
    assert _draw(rect[0, :n][keep], kp[0, :n][keep], seg[0, :n][keep], rec["size"], False, None) == rec["sha256"]
    assert _draw(rect[0, :n][keep], kp[0, :n][keep], seg[0, :n][keep], rec["size"], True,
                 packed.part_box[0, :n].cpu().numpy()[keep]) == rec["sha256_visbbox"]


@pytest.mark.gpu
def test_gpu_rt_test_inference_returns_the_references_image():
    """rt_test.inference (rt_test.py:87-147) with a stand-in network that returns a golden head tensor: the PIL image
    equals the one the reference's draw_humans drew for the same humans."""
    import torch
    from pytorch_pose_proposal_network_b200 import config as pcfg, rt_test
    name = next(n for n in sorted(CASES) if len(CASES[n]["kept"]) == load_case(n)[2]["part_cell"].shape[0])
    rec = CASES[name]
    g, out, fx = load_case(name)

    class Net(torch.nn.Module):
        keypoint_names, edges, insize = pcfg.KEYPOINT_NAMES, pcfg.EDGES, (g.inW, g.inH)

        def forward(self, x):
            return torch.from_numpy(out[None]).to(x.device)

    frame = np.zeros((g.inH, g.inW, 3), np.uint8)
    img = rt_test.inference(frame, Net(), (g.W, g.H), (g.sW, g.sH), image_size=g.inW)
    assert hashlib.sha256(np.asarray(img).tobytes()).hexdigest() == rec["sha256"]
    humans, scores = rt_test.inference(frame, Net(), (g.W, g.H), (g.sW, g.sH), image_size=g.inW, return_humans=True)
    assert len(humans) == fx["part_cell"].shape[0] and list(humans[0].keys())[0] == 0
