"""Import the reference's own parser, unmodified, for pinning the oracle.  TEST INFRASTRUCTURE.

Works only where ``/root/reference`` exists (the build container — never the GPU box).  The
reference's ``datatest.py`` imports plotting / augmentation packages at module level that
this image does not have (datatest.py:11-14,24,37-38 and, through ``dataset.py``/``aug.py``/
``eval_helpers.py``, skimage, imgaug, shapely); none of them is touched by the parser
functions, so empty stand-in modules are registered for them before the import.  Nothing is
copied: the functions executed are the reference's files where they lie.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_DIR = os.environ.get("PPN_REFERENCE_DIR", "/root/reference")

_MISSING = [
    "skimage", "skimage.io", "skimage.transform", "skimage.color",
    "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
    "imgaug", "imgaug.augmenters", "torchsummary", "shapely", "shapely.geometry",
]


class _Absent(types.ModuleType):
    """A module whose every attribute is another such module and which can be called."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Absent(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *args, **kwargs):
        return None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "datatest.py"))


_datatest = None


def load():
    """Return the reference's ``datatest`` module (cached)."""
    global _datatest
    if _datatest is not None:
        return _datatest
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    for name in _MISSING:
        try:
            importlib.import_module(name)
        except Exception:
            mod = _Absent(name)
            mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
            mod.__path__ = []
            sys.modules[name] = mod
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    # the reference's flat module names ('config', 'utils', 'dataset', 'aug') must resolve
    # to ITS files; drop same-named modules imported from elsewhere first
    for flat in ("config", "utils", "dataset", "aug", "datatest", "eval_helpers", "evaluateAP"):
        m = sys.modules.get(flat)
        if m is not None and not str(getattr(m, "__file__", "")).startswith(REFERENCE_DIR):
            del sys.modules[flat]
    import logging
    level = logging.getLogger().level
    _datatest = importlib.import_module("datatest")
    logging.getLogger().setLevel(max(level, logging.WARNING))
    return _datatest


def configure(g):
    """Patch the module globals the reference's functions read (datatest.py:53-60) and its
    star-imported DIRECTED_GRAPHS (config.py:75-80) to the geometry ``g`` (oracle Geometry)."""
    dt = load()
    dt.insize = (g.inW, g.inH)
    dt.outsize = (g.W, g.H)
    dt.local_grid_size = (g.sW, g.sH)
    dt.gridsize = (int(g.inW / g.W), int(g.inH / g.H))
    dt.DIRECTED_GRAPHS = [[list(eis), list(ts)] for eis, ts in g.graphs]
    return dt


def reference_parse(out, g):
    """Run rt_test.py:109-133's host steps and the reference parser on one image [C,H,W]."""
    dt = configure(g)
    K = g.K
    resp, conf, x, y, w, h = (out[i * K:(i + 1) * K] for i in range(6))
    e = out[6 * K:].reshape(g.E, g.sH, g.sW, g.H, g.W)
    return dt.get_humans_by_feature(resp * conf, x, y, w, h, e,
                                    detection_thresh=g.det_thresh, min_num_keypoints=g.min_kp)


def reference_draw(humans, size, visbbox=False):
    """The reference's own ``draw_humans`` (datatest.py:162-232) on a black ``size`` x ``size`` RGB image, as
    rt_test.py:138-145 calls it (its own KEYPOINT_NAMES / EDGES: the 18-part skeleton).  -> uint8 [size, size, 3]."""
    import numpy as np
    from PIL import Image
    dt = load()
    img = dt.draw_humans(keypoint_names=dt.KEYPOINT_NAMES, edges=dt.EDGES, pil_image=Image.new("RGB", (size, size)),
                         humans=humans, visbbox=visbbox, gridOn=False)
    return np.asarray(img)


class _Captured(Exception):
    def __init__(self, gt, pr):
        self.gt, self.pr = gt, pr


def reference_pred_frames(fnames, humans_list, scores_list):
    """The prediction frames the reference's own ``evaluation`` builds (datatest.py:298-348), captured
    at the point where it hands them to the poseval code (``eval_helpers.load_data``, :350)."""
    dt = load()

    def grab(gt, pr):
        raise _Captured(gt, pr)

    saved = dt.eval_helpers.load_data
    dt.eval_helpers.load_data = grab
    try:
        n = len(fnames)
        dt.evaluation([list(fnames), [[] for _ in range(n)], list(humans_list), list(scores_list),
                       [[] for _ in range(n)], [[] for _ in range(n)], [[] for _ in range(n)]])
    except _Captured as c:
        return c.pr
    finally:
        dt.eval_helpers.load_data = saved
    raise RuntimeError("the reference did not reach eval_helpers.load_data")


def reference_encode_targets(samples, insize, outsize, local_grid_size, keypoint_names, edges):
    """Run the reference's own ``KeypointsDataset.__getitem__`` (dataset.py:70-198), unmodified, on
    prepared samples and return its target lists.

    ``samples``: list of dicts as the reference's transform pipeline hands them to the encoder
    (aug.py:138-160): 'keypoints' torch fp32 [n, K-1, 2], 'bbox' torch float64 [n, 4] (cx, cy, w, h),
    'is_visible' list of bool arrays [K-1], 'size' list of floats, 'image' anything.  The dataset object
    is created without its __init__ (which only reads the annotation JSON and uses the removed
    ``np.bool``, dataset.py:53); image reading is stubbed; the ``transform`` hook returns the prepared
    sample.  Everything after "# Encode samples" (dataset.py:89) is the reference's code as it lies.
    """
    load()
    import numpy as np
    ds_mod = importlib.import_module("dataset")
    ds = object.__new__(ds_mod.KeypointsDataset)
    names = [f"{i:04d}.jpg" for i in range(len(samples))]
    ds.filename_list = names
    ds.root_dir = ""
    ds.draw = False
    ds.data = {n: ([np.zeros((len(keypoint_names) - 1, 2), np.float32)], [], [], []) for n in names}
    ds.insize, ds.outsize = tuple(insize), tuple(outsize)
    ds.keypoint_names, ds.local_grid_size, ds.edges = list(keypoint_names), tuple(local_grid_size), [list(e) for e in edges]
    ds.inW, ds.inH = insize
    ds.outW, ds.outH = outsize
    ds.gridW, ds.gridH = int(ds.inW / ds.outW), int(ds.inH / ds.outH)
    by_name = dict(zip(names, samples))
    current = {}
    ds.transform = lambda sample: by_name[current["name"]]
    saved = ds_mod.io.imread
    ds_mod.io.imread = lambda path: np.zeros((2, 2, 3), np.uint8)
    out = []
    try:
        for i, n in enumerate(names):
            current["name"] = n
            item = ds[i]
            out.append([t.numpy().copy() for t in item[1:]])      # drop the image
    finally:
        ds_mod.io.imread = saved
    return out          # per sample: [delta, weight, weight_ij, tx, ty, tx_half, ty_half, tw, th, te]
