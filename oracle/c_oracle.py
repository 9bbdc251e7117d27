"""ctypes binding of oracle/libppn_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY.

Builds the library on first use with oracle/Makefile (gcc only).  See ppn_oracle.c for the
reference lines restated; see ppn_oracle.py for the numpy twin.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libppn_oracle.so")


class _Geom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("K", "E", "inW", "inH", "W", "H", "sW", "sH", "off_h", "off_w", "n_chains")] + [
        ("chain_off", C.POINTER(C.c_int32)), ("chain_limb", C.POINTER(C.c_int32)), ("chain_part", C.POINTER(C.c_int32)),
        ("det_thr", C.c_float), ("nms_thr", C.c_float), ("min_kp", C.c_int32)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ppn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libppn_oracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)
        _lib.ppn_oracle_limb_argmax.argtypes = [f32p, C.POINTER(_Geom), i32p]
        _lib.ppn_oracle_window_argmax.argtypes = [f32p, C.POINTER(_Geom), C.c_int, C.c_int]
        _lib.ppn_oracle_nms.argtypes = [f32p, f32p, C.c_int, C.c_float, C.c_int, i32p]
        _lib.ppn_oracle_parse_image.argtypes = [f32p, C.POINTER(_Geom), i32p, i32p, i32p, i32p, i32p, f32p, f32p, i32p]
        _lib.ppn_oracle_parse_batch.argtypes = [f32p, C.c_int, C.POINTER(_Geom), C.c_int, i32p, i32p, i32p, i32p, i32p, f32p, f32p]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class CGeometry:
    """Keeps the chain arrays alive next to the C struct."""

    def __init__(self, g):
        off, limbs, parts = [0], [], []
        for eis, ts in g.graphs:
            limbs += list(eis); parts += list(ts); off.append(len(limbs))
        self.off = np.asarray(off, np.int32)
        self.limbs = np.asarray(limbs if limbs else [0], np.int32)
        self.parts = np.asarray(parts if parts else [0], np.int32)
        self.g = g
        self.c = _Geom(g.K, g.E, g.inW, g.inH, g.W, g.H, g.sW, g.sH, g.oh, g.ow, len(g.graphs),
                       _p(self.off, C.c_int32), _p(self.limbs, C.c_int32), _p(self.parts, C.c_int32),
                       float(np.float32(g.det_thresh)), float(np.float32(g.nms_thresh)), g.min_kp)


def limb_argmax(out: np.ndarray, g) -> np.ndarray:
    cg = CGeometry(g)
    out = np.ascontiguousarray(out, np.float32)
    amax = np.empty((g.E, g.H, g.W), np.int32)
    rc = lib().ppn_oracle_limb_argmax(_p(out, C.c_float), C.byref(cg.c), _p(amax, C.c_int32))
    assert rc == 0
    return amax


def window_argmax(out: np.ndarray, g, ei: int, cell: int) -> int:
    cg = CGeometry(g)
    out = np.ascontiguousarray(out, np.float32)
    return int(lib().ppn_oracle_window_argmax(_p(out, C.c_float), C.byref(cg.c), ei, cell))


def nms(bbox, thresh, score=None, limit=None) -> np.ndarray:
    bbox = np.ascontiguousarray(bbox, np.float32).reshape(-1, 4)
    n = bbox.shape[0]
    keep = np.empty(max(n, 1), np.int32)
    sp = None
    if score is not None:
        score = np.ascontiguousarray(score, np.float32)
        sp = _p(score, C.c_float)
    m = lib().ppn_oracle_nms(_p(bbox, C.c_float), sp, n, float(np.float32(thresh)), int(limit or 0), _p(keep, C.c_int32))
    assert m >= 0
    return keep[:m].copy()


def parse_batch(head: np.ndarray, g, n_threads: int = 1):
    """head [B,C,H,W] fp32 -> dict of packed arrays (same layout as the CUDA path writes)."""
    head = np.ascontiguousarray(head, np.float32)
    B = head.shape[0]
    assert head.shape[1:] == (g.C, g.H, g.W), head.shape
    HW, K = g.H * g.W, g.K
    cg = CGeometry(g)
    r = dict(counts=np.zeros((B, 3), np.int32), cand_cell=np.zeros((B, HW), np.int32),
             keep_idx=np.zeros((B, HW), np.int32), root_cell=np.zeros((B, HW), np.int32),
             part_cell=np.full((B, HW, K), -1, np.int32), part_score=np.zeros((B, HW, K), np.float32),
             part_box=np.zeros((B, HW, K, 4), np.float32))
    rc = lib().ppn_oracle_parse_batch(_p(head, C.c_float), B, C.byref(cg.c), n_threads,
                                      _p(r["counts"], C.c_int32), _p(r["cand_cell"], C.c_int32),
                                      _p(r["keep_idx"], C.c_int32), _p(r["root_cell"], C.c_int32),
                                      _p(r["part_cell"], C.c_int32), _p(r["part_score"], C.c_float),
                                      _p(r["part_box"], C.c_float))
    assert rc == 0
    return r
