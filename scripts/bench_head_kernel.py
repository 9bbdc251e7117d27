"""The fused head kernel alone (ppn_head_gemm_argmax_opt: clear, [pack,] GEMM + epilogue, maxima -> map), without the parse,
optionally with the epilogue skipped (head.dry: results invalid) — what the operand stream and the MMAs alone take.

    python scripts/bench_head_kernel.py [--configs cfg2,native] [--subs 3,4] [--operands f16,tf32]
"""
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
ap.add_argument("--configs", default="cfg2,native")
ap.add_argument("--subs", default="3,4")
ap.add_argument("--operands", default="f16,tf32")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--acc", type=int, default=256, help="16-bit path: channels per accumulator (128 x 4 or 256 x 2)")
args = ap.parse_args()
_lib.tune(head_acc=args.acc)


def timed(fn, iters):
    for i in range(3):
        fn(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for name in args.configs.split(","):
    cfg = PRESETS[name]()
    B, Cin = BATCH[name], 512
    gen = torch.Generator(device="cuda").manual_seed(5)
    feats = [torch.randn(B, Cin, cfg.H, cfg.W, device="cuda", generator=gen) for _ in range(3)]
    weight = (torch.randn(cfg.C, Cin, device="cuda", generator=gen) * (2.0 / (1.01 * Cin)) ** 0.5).contiguous()
    bias = torch.randn(cfg.C, device="cuda", generator=gen) * 0.5
    parser = PoseParser(cfg)
    flops = 2.0 * B * cfg.HW * cfg.C * Cin
    print(f"{name}: B={B} C={cfg.C} grid {cfg.H}x{cfg.W}: {flops / 1e9:.1f} GFLOP")
    for op in args.operands.split(","):
        ins = feats if op == "tf32" else [f.to({"f16": torch.float16, "bf16": torch.bfloat16}[op]).contiguous(memory_format=torch.channels_last) for f in feats]
        for subs in [int(s) for s in args.subs.split(",")]:
            row = []
            for dry in (0, 1, 2):
                _lib.tune(head_subs=subs, head_dry=dry)
                t = timed(lambda i: parser.head_gemm_argmax(ins[i % 3], weight, bias, operand=op), args.iters)
                row.append(t)
            _lib.tune(head_dry=0)
            print(f"   {op:5s} subs={subs}: kernel chain {row[0]:8.1f} us ({flops / row[0] / 1e6:6.1f} TFLOP/s)   epilogue skipped {row[1]:8.1f} us ({flops / row[1] / 1e6:6.1f} TFLOP/s)   1/8 of the MMAs (16-bit path) {row[2]:8.1f} us")
