"""The fused network head (ppn_head_parse: tcgen05 1x1-conv GEMM + sigmoid + arg-max epilogue, then the fused parse)
next to what it replaces: cuDNN's 1x1 convolution in TF32 + sigmoid (writing the fp32 head tensor) + ppn_parse reading it.

    python scripts/bench_head.py [--configs cfg2,native] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_pose_proposal_network_b200.config import PRESETS  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402

BATCH = {"cfg2": 512, "cfg3": 1024, "cfg4": 256, "native": 64}
ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="cfg2,native")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--cin", type=int, default=512)
ap.add_argument("--batch", type=int, default=0, help="images per call (default: the preset's throughput batch)")
ap.add_argument("--dry", action="store_true", help="epilogue skipped (results invalid): what the operand stream and the MMAs alone take")
ap.add_argument("--subs", type=int, default=0, help="epilogue warps per TMEM lane quadrant (tune key head.subs), 0 = library default")
args = ap.parse_args()
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
bf16 = json.load(open(peaks))["bf16_tflops"] if os.path.exists(peaks) else 1590.0
tf32_peak = bf16 / 2                                     # TF32 runs at half the bf16 rate
if args.subs:
    from pytorch_pose_proposal_network_b200 import _lib
    _lib.tune(head_subs=args.subs)
    print(f"# head.subs = {args.subs}")
if args.dry:
    from pytorch_pose_proposal_network_b200 import _lib
    _lib.tune(head_dry=1)
    print("# head.dry = 1: the epilogue releases the accumulators unread; times are the operand stream + MMAs only")
torch.backends.cudnn.allow_tf32 = True                   # PyTorch's default: the reference's conv3 runs in TF32


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
    return a.elapsed_time(b) / iters


for name in args.configs.split(","):
    cfg = PRESETS[name]()
    B, Cin = args.batch or BATCH[name], args.cin
    gen = torch.Generator(device="cuda").manual_seed(5)
    feats = [torch.randn(B, Cin, cfg.H, cfg.W, device="cuda", generator=gen) for _ in range(3)]
    weight = (torch.randn(cfg.C, Cin, device="cuda", generator=gen) * (2.0 / (1.01 * Cin)) ** 0.5).contiguous()
    bias = torch.randn(cfg.C, device="cuda", generator=gen) * 0.5
    bias[:2 * cfg.K] += 1.0
    w4 = weight[:, :, None, None].contiguous()
    parser = PoseParser(cfg)
    outs = [parser.alloc_output(B) for _ in range(2)]
    flops = 2.0 * B * cfg.HW * cfg.C * Cin

    def fused(i):
        parser.parse_features(feats[i % 3], weight, bias, out=outs[i % 2])

    def unfused(i):
        head = torch.sigmoid(torch.nn.functional.conv2d(feats[i % 3], w4, bias))
        parser.parse(head, out=outs[i % 2])

    def conv_only(i):
        torch.nn.functional.conv2d(feats[i % 3], w4, bias)

    t_f, t_u, t_c = timed(fused, args.iters), timed(unfused, args.iters), timed(conv_only, args.iters)
    if B <= 16:                                            # small batches are bound by the host's launches: also replay a CUDA graph
        cap = parser.capture_features(feats[0], weight, bias, out=outs[0])
        cap16 = parser.capture_features(feats[1].half().contiguous(memory_format=torch.channels_last), weight, bias, out=outs[1], operand="f16") \
            if Cin % 64 == 0 and Cin <= 512 else None
        t_g = timed(lambda i: cap.replay(), args.iters)
        t_g16 = timed(lambda i: cap16.replay(), args.iters) if cap16 else float("nan")
        print(f"{name}: B={B}: fused head+parse replayed from a CUDA graph: TF32 {t_g * 1e3:.1f} us, f16 channels_last {t_g16 * 1e3:.1f} us per call")
    print(f"{name}: B={B} Cin={Cin} C={cfg.C} grid {cfg.H}x{cfg.W}: {flops / 1e9:.1f} GFLOP per batch")
    print(f"   TF32 operands (fp32 NCHW activations read in place)")
    print(f"      fused head+parse   {t_f * 1e3:8.1f} us  {B / t_f / 1e3:8.1f} k img/s  {flops / t_f / 1e9:7.1f} TFLOP/s = {flops / t_f / 1e9 / tf32_peak:.3f} of the TF32 peak "
          f"({tf32_peak:.0f} TF/s = measured bf16 {bf16:.0f} / 2)")
    print(f"      cuDNN conv (TF32) + sigmoid + ppn_parse   {t_u * 1e3:8.1f} us  {B / t_u / 1e3:8.1f} k img/s   (conv alone {t_c * 1e3:.1f} us = {flops / t_c / 1e9:.1f} TFLOP/s)")
    if Cin % 64 == 0 and Cin <= 512:
        for op, dt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
            feats_cl = [f.to(dt).contiguous(memory_format=torch.channels_last) for f in feats]
            w4h, bh = w4.to(dt).contiguous(memory_format=torch.channels_last), bias.to(dt)

            def fused16(i):
                parser.parse_features(feats[i % 3], weight, bias, out=outs[i % 2], operand=op)

            def fused16_cl(i):
                parser.parse_features(feats_cl[i % 3], weight, bias, out=outs[i % 2], operand=op)

            def conv16(i):
                torch.nn.functional.conv2d(feats_cl[i % 3], w4h, bh)

            t1, t2, t3 = timed(fused16, args.iters), timed(fused16_cl, args.iters), timed(conv16, args.iters)
            print(f"   {op} operands, fp32 accumulation (peak {bf16:.0f} TF/s measured)")
            print(f"      fused head+parse, fp32 NCHW in (pack pre-pass)   {t1 * 1e3:8.1f} us  {B / t1 / 1e3:8.1f} k img/s  {flops / t1 / 1e9:7.1f} TFLOP/s = {flops / t1 / 1e9 / bf16:.3f} of the peak")
            print(f"      fused head+parse, channels_last {op} in (in place) {t2 * 1e3:8.1f} us  {B / t2 / 1e3:8.1f} k img/s  {flops / t2 / 1e9:7.1f} TFLOP/s = {flops / t2 / 1e9 / bf16:.3f} of the peak")
            print(f"      cuDNN conv alone ({op}, channels_last)            {t3 * 1e3:8.1f} us  {flops / t3 / 1e9:7.1f} TFLOP/s")
            del feats_cl
    del feats, outs, parser
    torch.cuda.empty_cache()
