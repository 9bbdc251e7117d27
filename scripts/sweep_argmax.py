"""Sweep the limb arg-max kernel's knobs on one GPU and print achieved GB/s per setting.

    python scripts/sweep_argmax.py [--presets cfg2,cfg3,cfg4,native] [--iters 30] > gpurun_out/sweep.txt
"""
import argparse
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_pose_proposal_network_b200 import _lib  # noqa: E402
from pytorch_pose_proposal_network_b200.config import PRESETS  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402

BATCH = {"cfg2": 512, "cfg3": 1024, "cfg4": 256, "native": 64}


def time_setting(parser, bufs, iters):
    for i in range(3):
        parser.limb_argmax(bufs[i % len(bufs)])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        parser.limb_argmax(bufs[i % len(bufs)])
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--presets", default="cfg2,cfg3,cfg4,native")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--tail-opt", type=int, default=1)
    ap.add_argument("--head-dtype", default="f32", choices=["f32", "f16", "bf16"])
    args = ap.parse_args()
    peak = 6553.0
    _lib.tune(argmax_tail_opt=args.tail_opt)
    for name in args.presets.split(","):
        cfg = PRESETS[name]()
        B = BATCH[name]
        nbuf = 3 if B * cfg.bytes_per_image < (1 << 30) else 2
        dt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[args.head_dtype]
        if dt != torch.float32:
            nbuf = 4                                      # half the bytes per batch: keep the rotation larger than L2
        bufs = [torch.rand(B, cfg.C, cfg.H, cfg.W, device="cuda").to(dt) for _ in range(nbuf)]
        parser = PoseParser(cfg)
        limb_bytes = B * cfg.E * cfg.S * cfg.HW * bufs[0].element_size()
        rows = []
        grid = list(itertools.product([0], [16384, 32768, 49152, 65536], [3, 4, 5, 6], [192, 320, 512], [1, 2], [0, 1]))
        grid += list(itertools.product([1], [32768], [5], [128, 192], [1], [0]))
        if dt != torch.float32:                           # the 16-bit ring always splits matrices; more thread counts instead
            grid = list(itertools.product([0], [16384, 32768, 49152], [3, 4, 5, 6], [160, 320, 480, 576, 768], [1, 2], [1]))
        if args.quick:
            grid = grid[::7]
        for variant, sb, st, th, ctas, split in grid:
            if variant == 0 and sb * st * ctas > 215 * 1024:
                continue
            _lib.tune(argmax_variant=variant, argmax_stage_bytes=sb, argmax_stages=st, argmax_threads=th, argmax_ctas_per_sm=ctas, argmax_split=split,
                      argmax16_stage_bytes=sb, argmax16_threads=th)
            try:
                med, best = time_setting(parser, bufs, args.iters)
            except Exception as e:
                print(name, variant, sb, st, th, ctas, split, "ERR", e)
                torch.cuda.synchronize()
                continue
            gbs = limb_bytes / (med * 1e-3) / 1e9
            rows.append((gbs, variant, sb, st, th, ctas, split, med, best))
        rows.sort(reverse=True)
        print(f"== {name}: B={B}, limb block {limb_bytes / 1e6:.0f} MB; 100% of {peak} GB/s = {limb_bytes / peak / 1e3:.1f} us")
        for gbs, variant, sb, st, th, ctas, split, med, best in rows[:14] + rows[-2:]:
            print(f"  {gbs:7.0f} GB/s {gbs / peak * 100:5.1f}%  variant={variant} stage_bytes={sb} stages={st} threads={th} ctas/sm={ctas} split={split}  "
                  f"median {med * 1e3:.1f} us  best {best * 1e3:.1f} us")
        print(json.dumps({"preset": name, "rows": rows}))
        del bufs, parser
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
