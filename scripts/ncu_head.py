"""One launch of the fused head kernels for an ncu capture (cfg2 shape, 128 images by default):

    ncu --set full --clock-control none --import-source on -k regex:'head_gemm|head_pack' -o gpurun_out/head \\
        python scripts/ncu_head.py [operand] [config] [B]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_pose_proposal_network_b200.config import PRESETS  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402

operand = sys.argv[1] if len(sys.argv) > 1 else "f16"
cfg = PRESETS[sys.argv[2] if len(sys.argv) > 2 else "cfg2"]()
B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
gen = torch.Generator(device="cuda").manual_seed(11)
feat = torch.randn(B, 512, cfg.H, cfg.W, device="cuda", generator=gen)
weight = torch.randn(cfg.C, 512, device="cuda", generator=gen) * 0.06
bias = torch.randn(cfg.C, device="cuda", generator=gen) * 0.5
bias[:2 * cfg.K] += 1.0
if os.environ.get("HEAD_SUBS"):
    from pytorch_pose_proposal_network_b200 import _lib
    _lib.tune(head_subs=int(os.environ["HEAD_SUBS"]))
p = PoseParser(cfg)
out = p.parse_features(feat, weight, bias, operand=operand)
torch.cuda.synchronize()
print("humans/image", float(out.count.float().mean()))
