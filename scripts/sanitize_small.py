"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel variant once.

    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle, ppn_oracle as O, synth  # noqa: E402
from pytorch_pose_proposal_network_b200 import _lib  # noqa: E402
from pytorch_pose_proposal_network_b200.config import PPNConfig  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402


def check(cfg, B, dist, **tune):
    g = O.Geometry.of(cfg)
    head = synth.make_head(g, dist, seed=3, B=B)
    ref = c_oracle.parse_batch(head, g)
    _lib.tune(**tune)
    p = PoseParser(cfg)
    dev = torch.from_numpy(head).cuda()
    for i in range(3):
        out = p.parse(dev, input_complete=(i > 0))
    a = out.numpy()
    assert np.array_equal(a["count"], ref["counts"][:, 2])
    for b in range(B):
        n = int(a["count"][b])
        assert np.array_equal(a["part_cell"][b, :n], ref["part_cell"][b, :n])
    buf = p.pack(out, 64)
    torch.cuda.synchronize()
    print("ok", cfg.H, cfg.W, dist, tune)


small = PPNConfig.mpii16()
check(small, 3, "U")
check(small, 3, "D", argmax_split=0)
check(small, 2, "U", argmax_variant=1)
check(small, 2, "U", argmax_variant=0, argmax_split=-1, parse_overlap=1)
check(small, 2, "U", parse_overlap=0, parse_stage_all=2)
check(PPNConfig.mpii16(insize=(208, 208), outsize=(13, 13)), 2, "U", parse_overlap=2, parse_stage_all=-1)   # generic arg-max, no TMA staging
check(PPNConfig.highres(), 1, "U")
print("done")
