"""Device-side timeline of consecutive overlapped ppn_parse calls (no profiler needed): every kernel records its first
CTA's start, its last CTA's end and the moment it got past its dependency wait in %globaltimer ns (ppn_timeline).

    python scripts/timeline.py [--config cfg2] [--steps 6] [--tune parse.persist=2]
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
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--warm", type=int, default=20)
ap.add_argument("--tune", action="append", default=[])
args = ap.parse_args()
for kv in args.tune:
    k, v = kv.split("=")
    _lib.tune(**{k.replace(".", "_"): int(v)})
cfg = PRESETS[args.config]()
B = BATCH[args.config]
parser = PoseParser(cfg)
gen = torch.Generator(device="cuda").manual_seed(3)
bufs = []
for _ in range(3):
    t = torch.rand(B, cfg.C, cfg.H, cfg.W, device="cuda", generator=gen)
    if args.config == "cfg3":
        t[:, :2 * cfg.K] = 0.4 + 0.6 * t[:, :2 * cfg.K]
        t[:, 4 * cfg.K:6 * cfg.K] *= 0.08
    bufs.append(t)
outs = [parser.alloc_output(B) for _ in range(2)]
for i in range(args.warm):
    parser.parse(bufs[i % 3], out=outs[i % 2], input_complete=True)
torch.cuda.synchronize()
n_k = parser.parse_plan(B)["launches"]
rec = torch.zeros(args.steps * n_k, 4, dtype=torch.int64, device="cuda")
rec[:, 0] = rec[:, 2] = torch.iinfo(torch.int64).max
_lib.check(_lib.lib().ppn_timeline(rec.data_ptr(), rec.shape[0]), "ppn_timeline")
for i in range(args.steps):
    parser.parse(bufs[i % 3], out=outs[i % 2], input_complete=True)
torch.cuda.synchronize()
_lib.check(_lib.lib().ppn_timeline(None, 0), "ppn_timeline")
r = rec.cpu().numpy()
t0 = int(r[:, 0].min())
names = {2: ["argmax", "parse_fused"], 3: ["decode_nms", "argmax", "tree_parse"]}.get(n_k, [f"k{j}" for j in range(n_k)])
print(f"# {args.config} B={B} tune={args.tune}: kernel, call, first CTA start, past its wait, last CTA end (us since the first start)")
for i in range(args.steps):
    for j in range(n_k):
        s, e, w, _ = (int(v) for v in r[i * n_k + j])
        waited = f"{(w - t0) / 1e3:8.1f}" if w < (1 << 62) else "       -"
        print(f"{names[j % len(names)]:12s} call {i}  start {(s - t0) / 1e3:8.1f}  waited {waited}  end {(e - t0) / 1e3:8.1f}  ({(e - s) / 1e3:6.1f} us)")
ends = [int(r[i * n_k + n_k - 1, 1]) for i in range(args.steps)]
print("# step period (last kernel's end to end): " + ", ".join(f"{(b - a) / 1e3:.1f}" for a, b in zip(ends, ends[1:])) + " us")
