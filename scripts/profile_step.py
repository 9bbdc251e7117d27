"""A few back-to-back parses of one config, for ncu (no timing, no CPU baseline).

    python scripts/profile_step.py [--config cfg2] [--steps 4] [--overlap 0]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_pose_proposal_network_b200 import _lib  # noqa: E402
from pytorch_pose_proposal_network_b200.config import PRESETS  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402

BATCH = {"cfg1": 1, "cfg2": 512, "cfg3": 1024, "cfg4": 256, "native": 64}

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--overlap", type=int, default=0)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--dense", action="store_true")
ap.add_argument("--tune", action="append", default=[])
ap.add_argument("--head-dtype", default="f32", choices=["f32", "f16", "bf16"])
a = ap.parse_args()
_lib.tune(parse_overlap=a.overlap)
for kv in a.tune:
    k, v = kv.split("=")
    _lib.tune(**{k.replace(".", "_"): int(v)})
cfg = PRESETS[a.config]()
B = a.batch or BATCH[a.config]
bufs = [torch.rand(B, cfg.C, cfg.H, cfg.W, device="cuda") for _ in range(2)]
dt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[a.head_dtype]
if a.dense or a.config == "cfg3":
    for t in bufs:
        t[:, :2 * cfg.K] = 0.4 + 0.6 * t[:, :2 * cfg.K]
        t[:, 4 * cfg.K:6 * cfg.K] *= 0.08
bufs = [t.to(dt) for t in bufs]
p = PoseParser(cfg)
for i in range(a.steps):
    out = p.parse(bufs[i % 2])
torch.cuda.synchronize()
print("humans/image", float(out.count.float().mean()))
