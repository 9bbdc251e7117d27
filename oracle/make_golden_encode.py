"""Golden vectors of the reference's training-target encoder.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_encode        (build container only: needs /root/reference)

Runs the reference's own ``KeypointsDataset.__getitem__`` (dataset.py:70-198, through
``oracle/ref_live.reference_encode_targets``) on seeded synthetic annotation sets, asserts that the
numpy restatement ``oracle/encode_gt.encode_targets`` reproduces every output bit for bit, and writes
the reference's outputs to ``tests/golden/encode/*.npz``: the eight [K,H,W] grids in full, ``te`` and
``weight_ij`` as the flat indices of their ones (their other elements are 0 and float32(0.0005)).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import encode_gt, ref_live  # noqa: E402
from pytorch_pose_proposal_network_b200 import config as pcfg  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "encode")
NAMES = ("delta", "weight", "weight_ij", "tx", "ty", "tx_half", "ty_half", "tw", "th", "te")

CASES = {
    # name: (insize, outsize, window, K, seed, people per image)
    "enc_cfg2like": ((384, 384), (12, 12), (9, 9), 18, 101, [3, 0, 8, 1, 5, 12]),
    "enc_native": ((384, 384), (24, 24), (21, 21), 18, 102, [4, 7, 0, 2]),
    "enc_cfg4like": ((768, 768), (24, 24), (11, 11), 18, 103, [6, 1, 9]),
    "enc_k16_rect": ((384, 256), (12, 8), (5, 5), 16, 104, [2, 5, 0, 3, 3]),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (insize, outsize, window, K, seed, counts) in CASES.items():
        rng = np.random.default_rng(seed)
        names = pcfg.KEYPOINT_NAMES if K == 18 else pcfg.KEYPOINT_NAMES_16
        edges = pcfg.EDGES if K == 18 else pcfg.EDGES_16
        samples, raw = [], []
        for n in counts:
            kp, bb, vis, size = encode_gt.random_people(rng, n, K, insize)
            raw.append((kp, bb, vis, size))
            kp_t = kp if n else np.zeros((1, K - 1, 2), np.float32)            # aug.py:115-116 keeps one zero person
            samples.append(dict(image=None, keypoints=torch.from_numpy(kp_t), bbox=torch.from_numpy(np.asarray(bb).reshape(-1, 4)),
                                is_visible=vis, size=size))
        ref = ref_live.reference_encode_targets(samples, insize, outsize, window, names, edges)
        for (kp, bb, vis, size), r in zip(raw, ref):
            mine = encode_gt.encode_targets(kp, bb, vis, size, K, edges, insize, outsize, window)
            for nm, a, b in zip(NAMES, r, mine):
                assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, nm)
        fx = dict(meta=json.dumps(dict(insize=insize, outsize=outsize, window=window, K=K, seed=seed, counts=counts,
                                       source="reference dataset.KeypointsDataset.__getitem__")),
                  person_off=np.concatenate([[0], np.cumsum(counts)]).astype(np.int32),
                  keypoints=np.concatenate([r[0].reshape(-1, K - 1, 2) for r in raw]).astype(np.float32),
                  bbox=np.concatenate([np.asarray(r[1], np.float64).reshape(-1, 4) for r in raw]),
                  visible=np.concatenate([np.asarray(r[2], bool).reshape(-1, K - 1) for r in raw]).astype(np.uint8),
                  size=np.concatenate([np.asarray(r[3], np.float64).reshape(-1) for r in raw]))
        for i, nm in enumerate(NAMES):
            stack = np.stack([r[i] for r in ref])
            if nm in ("te", "weight_ij"):
                other = np.float32(0.0) if nm == "te" else np.float32(0.0005)
                ones = np.flatnonzero(stack.reshape(-1) == 1.0)
                assert np.all((stack.reshape(-1) == 1.0) | (stack.reshape(-1) == other)), nm
                fx[nm + "_ones"] = ones.astype(np.int64)
                fx[nm + "_shape"] = np.array(stack.shape, np.int64)
            else:
                fx[nm] = stack
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)
        print(name, {k: getattr(v, "shape", None) for k, v in fx.items() if k != "meta"})


if __name__ == "__main__":
    main()
