"""Seeded synthetic head tensors for parity tests and fixtures.  TEST INFRASTRUCTURE.

Distributions follow SURVEY.md §8(d); all are fp32 and generated with numpy's PCG64 so the
same (dist, seed, shape) gives the same bytes wherever the suite runs — fixtures store the
sha256 of the tensor they were made from and tests verify it before comparing.

  U  uniform [0,1)                                   "random tensor"
  R  0.5 + 0.5*z/(1+|z|), z standard normal          "random-init-like head output" (bell around 0.5)
  D  U with resp,conf in [0.4,1) and w,h * 0.08      "dense crowd": every cell is a root
  S  U with resp**8 and w,h * 0.3                    "sparse realistic"
"""
from __future__ import annotations

import hashlib

import numpy as np


def make_head(g, dist: str = "U", seed: int = 0, B: int = 1) -> np.ndarray:
    """[B, C, H, W] fp32 for geometry ``g`` (anything with K, C, H, W attributes)."""
    rng = np.random.default_rng(seed)
    shape = (B, g.C, g.H, g.W)
    K = g.K
    if dist == "R":
        # algebraic squashing instead of exp(): only +,*,/ so every platform rounds alike
        z = rng.standard_normal(shape, dtype=np.float32)
        out = np.float32(0.5) + np.float32(0.5) * (z / (np.float32(1) + np.abs(z)))
    else:
        out = rng.random(shape, dtype=np.float32)
        if dist == "D":
            out[:, 0:2 * K] = np.float32(0.4) + np.float32(0.6) * out[:, 0:2 * K]
            out[:, 4 * K:6 * K] *= np.float32(0.08)
        elif dist == "S":
            r2 = out[:, 0:K] * out[:, 0:K]          # x**8 by squaring: exactly rounded products
            r4 = r2 * r2
            out[:, 0:K] = r4 * r4
            out[:, 4 * K:6 * K] *= np.float32(0.3)
        elif dist != "U":
            raise ValueError(dist)
    return np.ascontiguousarray(out, np.float32)


def root_scores_distinct(out: np.ndarray, g, thr: float = 0.15) -> bool:
    """True when no image has two equal root scores above the threshold (tie order is the one
    thing the reference leaves undefined, datatest.py:139)."""
    K = g.K
    for img in out:
        d = (img[0] * img[K]).reshape(-1)
        d = d[d > np.float32(thr)]
        if np.unique(d).size != d.size:
            return False
    return True


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
