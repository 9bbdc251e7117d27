"""Load tests/golden/*.npz (made by oracle/make_golden.py from the live reference)."""
import glob
import json
import os

import numpy as np

from oracle import ppn_oracle as O, synth, encode_gt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not p.endswith("nms_cases.npz"))


def geometry_of(meta) -> O.Geometry:
    gm = meta["geometry"]
    return O.Geometry(K=gm["K"], E=gm["E"], inW=gm["inW"], inH=gm["inH"], W=gm["W"], H=gm["H"],
                      sW=gm["sW"], sH=gm["sH"], graphs=tuple((tuple(a), tuple(b)) for a, b in gm["graphs"]),
                      det_thresh=gm["det_thresh"], nms_thresh=gm["nms_thresh"], min_kp=gm["min_kp"])


def load_case(name):
    """-> (geometry, input [C,H,W] fp32, fixture dict).  The input is taken from the fixture when
    stored, otherwise regenerated from (dist, seed) and checked against the stored sha256."""
    fx = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    meta = json.loads(str(fx["meta"]))
    g = geometry_of(meta)
    if "input" in fx:
        out = fx["input"]
    elif meta["dist"] == "roundtrip":
        from pytorch_pose_proposal_network_b200 import config as pcfg
        people = json.loads(meta["people"])
        for p in people:
            p["points"] = {int(k): tuple(v) for k, v in p["points"].items()}
        out = encode_gt.encode_people(people, g, pcfg.EDGES if g.K == 18 else pcfg.EDGES_16)
        gridW = g.inW // g.W
        for i, person in enumerate(people):
            cx, cy = person["box"][:2]
            out[g.K, int(cy / gridW), int(cx / gridW)] = np.float32(1.0 - 0.03 * i)
    else:
        out = synth.make_head(g, meta["dist"], meta["seed"])[0]
    assert synth.digest(out) == str(fx["input_sha256"]), f"{name}: regenerated input differs from the fixture's"
    return g, out, fx


def load_nms_cases():
    z = np.load(os.path.join(GOLDEN, "nms_cases.npz"), allow_pickle=False)
    cases = {}
    for key in z.files:
        name, field = key.rsplit("/", 1)
        cases.setdefault(name, {})[field] = z[key]
    return cases
