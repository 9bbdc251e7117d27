"""cfg2 (512 images): one serial call, for an ncu capture of the fused parse kernel with source counters.

    ncu --set full --clock-control none --import-source on -k regex:'parse_fused' -c 1 -o gpurun_out/cfg2_k124 python scripts/ncu_cfg2_parse.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_pose_proposal_network_b200 import _lib  # noqa: E402
from pytorch_pose_proposal_network_b200.config import PRESETS  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402

gen = torch.Generator(device="cuda").manual_seed(11)
cfg = PRESETS["cfg2"]()
t = torch.rand(512, cfg.C, cfg.H, cfg.W, device="cuda", generator=gen)
_lib.tune(parse_overlap=0)
p = PoseParser(cfg)
out = p.parse(t)
torch.cuda.synchronize()
print("cfg2 humans/image", float(out.count.float().mean()))
