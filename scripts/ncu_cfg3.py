"""cfg3 (dense crowd, 1024 images of 16x16 cells, ~225 people each): the three kernels of the chain run serially, one call,
for an ncu capture of decode+NMS and the tree parse with source counters.

    ncu --set full --clock-control none --import-source on -k regex:'decode_nms|tree_parse' -o gpurun_out/cfg3 python scripts/ncu_cfg3.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_pose_proposal_network_b200 import _lib  # noqa: E402
from pytorch_pose_proposal_network_b200.config import PRESETS  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402

gen = torch.Generator(device="cuda").manual_seed(11)
cfg = PRESETS["cfg3"]()
B = 1024
t = torch.rand(B, cfg.C, cfg.H, cfg.W, device="cuda", generator=gen)
t[:, :2 * cfg.K] = 0.4 + 0.6 * t[:, :2 * cfg.K]
t[:, 4 * cfg.K:6 * cfg.K] *= 0.08
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    _lib.tune(**{k.replace(".", "_"): int(v)})
_lib.tune(parse_overlap=0)
p = PoseParser(cfg)
out = p.parse(t)
torch.cuda.synchronize()
print("cfg3 humans/image", float(out.count.float().mean()))
