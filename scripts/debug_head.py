"""Structured probes of the fused head kernel's GEMM (development aid): which (cell, k, channel) does each
operand element really land on?  python scripts/debug_head.py [preset] [Cin]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ppn_oracle as O
from pytorch_pose_proposal_network_b200.config import PRESETS
from pytorch_pose_proposal_network_b200.parser import PoseParser

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
Cin = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = PRESETS[name]()
parser = PoseParser(cfg)
B, HW, C = 2, cfg.HW, cfg.C
dev = "cuda"
np.set_printoptions(linewidth=200, precision=3, suppress=True)


def run(feat, weight):
    _, _, logits, _ = parser.head_gemm_argmax(feat.contiguous(), weight.contiguous(), None, emit=True)
    torch.cuda.synchronize()
    return logits.reshape(B, C, HW).cpu().numpy()


hw = torch.arange(HW, device=dev, dtype=torch.float32)
kk = torch.arange(Cin, device=dev, dtype=torch.float32)

# probe 1: feat = cell index, W one-hot at k0 -> logit[c, cell] == cell
for k0 in (0, 1, 9, 33):
    feat = hw.view(1, 1, HW).expand(B, Cin, HW).reshape(B, Cin, cfg.H, cfg.W).clone()
    w = torch.zeros(C, Cin, device=dev)
    w[:, k0] = 1.0
    out = run(feat, w)
    ok = np.array_equal(out, np.broadcast_to(np.arange(HW, dtype=np.float32), out.shape))
    print(f"probe1 k0={k0}: ok={ok}")
    if not ok:
        print(" b0 c0 cells[:40]:", out[0, 0, :40])
        print(" b0 c0 cells[120:]:", out[0, 0, 120:])
        print(" b1 c5 cells[:40]:", out[1, 5, :40])
        print(" b0 c300 cells[:40]:", out[0, 300, :40])

# probe 2: feat = k index, W[c, k] = 1 iff k == c % Cin -> logit[c, cell] == c % Cin
feat = kk.view(1, Cin, 1).expand(B, Cin, HW).reshape(B, Cin, cfg.H, cfg.W).clone()
w = torch.zeros(C, Cin, device=dev)
w[torch.arange(C, device=dev), torch.arange(C, device=dev) % Cin] = 1.0
out = run(feat, w)
want = (np.arange(C) % Cin).astype(np.float32)[None, :, None]
ok = np.array_equal(out, np.broadcast_to(want, out.shape))
print(f"probe2: ok={ok}")
if not ok:
    print(" b0 cell0 c[:70]:", out[0, :70, 0])
    print(" b0 cell77 c[:70]:", out[0, :70, 77])
    print(" b0 cell0 c[250:270]:", out[0, 250:270, 0])

# probe 3: feat = 1 at k = 0 only, W[c, 0] = c % 512 -> logit[c, cell] == c % 512
feat = torch.zeros(B, Cin, cfg.H, cfg.W, device=dev)
feat[:, 0] = 1.0
w = torch.zeros(C, Cin, device=dev)
w[:, 0] = (torch.arange(C, device=dev) % 512).float()
out = run(feat, w)
want = (np.arange(C) % 512).astype(np.float32)[None, :, None]
ok = np.array_equal(out, np.broadcast_to(want, out.shape))
print(f"probe3: ok={ok}")
if not ok:
    print(" b0 cell0 c[:40]:", out[0, :40, 0])
    print(" b0 cell0 c[250:290]:", out[0, 250:290, 0])
    print(" b1 cell100 c[1270:1311]:", out[1, 1270:, 100])

# probe 4: random, error map by (cell group, channel tile)
gen = torch.Generator(device=dev).manual_seed(1)
feat = torch.randn(B, Cin, cfg.H, cfg.W, device=dev, generator=gen)
w = torch.randn(C, Cin, device=dev, generator=gen) * 0.1
out = run(feat, w)
ref = torch.einsum("bkm,ck->bcm", feat.reshape(B, Cin, HW).double(), w.double()).float().cpu().numpy()
err = np.abs(out - ref)
print("probe4 max err", err.max(), "frac bad", float((err > 1e-2).mean()))
