"""One launch (or two) of every kernel that round 1 left without an ncu capture, each as it runs stand-alone:

  * cfg3 dense crowd, three-kernel chain run serially: limb arg-max ring, decode+NMS, tree parse
  * cfg2 with an fp16 head: the 16-bit ring (limb_argmax_tma_multi16_kernel)
  * the reference's native shape, one image: the cluster arg-max kernel and the fused parse kernel
  * the training-target encoder's sweep kernel (encode_sweep_kernel), cfg2 shape
  * the fused network head (head_gemm_argmax_kernel, tcgen05), cfg2 shape, 128 images

    ncu --set full --clock-control none --import-source on -k regex:'<names>' -o gpurun_out/prof python scripts/ncu_targets.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import encode_gt  # noqa: E402  (random annotations only)
from pytorch_pose_proposal_network_b200 import _lib  # noqa: E402
from pytorch_pose_proposal_network_b200.config import PRESETS  # noqa: E402
from pytorch_pose_proposal_network_b200.dataset import TargetEncoder, flatten_samples  # noqa: E402
from pytorch_pose_proposal_network_b200.parser import PoseParser  # noqa: E402

gen = torch.Generator(device="cuda").manual_seed(11)

# ---- cfg3, dense crowd, kernels one after the other ------------------------------------------------
cfg = PRESETS["cfg3"]()
B = 1024
t = torch.rand(B, cfg.C, cfg.H, cfg.W, device="cuda", generator=gen)
t[:, :2 * cfg.K] = 0.4 + 0.6 * t[:, :2 * cfg.K]
t[:, 4 * cfg.K:6 * cfg.K] *= 0.08
_lib.tune(parse_overlap=0)
p = PoseParser(cfg)
out = p.parse(t)
torch.cuda.synchronize()
print("cfg3 humans/image", float(out.count.float().mean()))
_lib.tune(parse_overlap=2)
del t, p, out
torch.cuda.empty_cache()

# ---- cfg2, fp16 head ------------------------------------------------------------------------------------
cfg = PRESETS["cfg2"]()
t16 = torch.rand(512, cfg.C, cfg.H, cfg.W, device="cuda", generator=gen).half()
p = PoseParser(cfg)
p.limb_argmax(t16)
torch.cuda.synchronize()

# ---- fused head at the cfg2 shape ------------------------------------------------------------------------
Bh, Cin = 128, 512
feat = torch.randn(Bh, Cin, cfg.H, cfg.W, device="cuda", generator=gen)
weight = torch.randn(cfg.C, Cin, device="cuda", generator=gen) * 0.06
bias = torch.randn(cfg.C, device="cuda", generator=gen) * 0.5
bias[:2 * cfg.K] += 1.0
out = p.parse_features(feat, weight, bias)
torch.cuda.synchronize()
print("head humans/image", float(out.count.float().mean()))
del t16, feat, weight, out

# ---- encoder sweep ---------------------------------------------------------------------------------------
rng = np.random.default_rng(0)
raw = [encode_gt.random_people(rng, 4, cfg.K, cfg.insize) for _ in range(512)]
samples = [dict(keypoints=kp, bbox=bb, is_visible=vis, size=size) for kp, bb, vis, size in raw]
enc = TargetEncoder(cfg)
flat = [torch.from_numpy(x).cuda() for x in flatten_samples(samples, cfg.K)]
enc.encode_flat(*flat, out=enc.alloc(512))
torch.cuda.synchronize()

# ---- native shape, one image: cluster arg-max + fused parse ------------------------------------------------
cfg = PRESETS["native"]()
one = torch.rand(1, cfg.C, cfg.H, cfg.W, device="cuda", generator=gen)
p1 = PoseParser(cfg)
out = p1.parse(one)
torch.cuda.synchronize()
print("native humans", int(out.count[0]))
