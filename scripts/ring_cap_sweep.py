"""How much shared memory does the arg-max ring need?  Times the kernel and the dry ring (same copies, no compares)
back to back over rotating inputs for several caps of the ring's shared memory.
    python scripts/ring_cap_sweep.py [--config cfg3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_pose_proposal_network_b200 import _lib  # noqa: E402
from pytorch_pose_proposal_network_b200.config import PRESETS  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402

BATCH = {"cfg2": 512, "cfg3": 1024, "cfg4": 256, "native": 64}
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg3")
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
cfg = PRESETS[args.config]()
B = BATCH[args.config]
parser = PoseParser(cfg)
bufs = [torch.rand(B, cfg.C, cfg.H, cfg.W, device="cuda") for _ in range(2 if B * cfg.C * cfg.HW * 4 > (1 << 29) else 3)]
limb = B * cfg.E * cfg.S * cfg.HW * 4


def timed(fn):
    for i in range(3):
        fn(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(args.iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / args.iters


print(f"# {args.config} B={B}: ring cap (bytes, 0 = all), arg-max kernel us / GB/s, dry ring us / GB/s")
for cap in (0, 160 << 10, 140 << 10, 120 << 10, 100 << 10, 80 << 10, 64 << 10):
    _lib.tune(argmax_smem_cap=cap)
    k = timed(lambda i: parser.limb_argmax_into(bufs[i % len(bufs)]))
    p = timed(lambda i: parser.limb_stream_probe(bufs[i % len(bufs)], cap))
    print(f"cap {cap:7d}: kernel {k * 1e3:7.1f} us {limb / k / 1e6:6.0f} GB/s   dry {p * 1e3:7.1f} us {limb / p / 1e6:6.0f} GB/s")
_lib.tune(argmax_smem_cap=0)
