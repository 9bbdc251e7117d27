"""Generate tests/golden/*.npz by running the REFERENCE's own functions.  TEST INFRASTRUCTURE.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

For every case this script
  1. builds a seeded input (oracle/synth.py, or oracle/encode_gt.py for the round-trip KAT),
  2. runs the unmodified reference parser on it (oracle/ref_live.py),
  3. asserts that the numpy restatement AND the C restatement reproduce the reference
     bit-for-bit (key order, box bits, score bits, NMS survivors),
  4. stores the reference's outputs, plus the cell ids the restatement found for them, in a
     small .npz.  Large inputs are not stored: the fixture keeps (dist, seed, shape) and the
     sha256 of the tensor, and tests regenerate and verify it.

The fixtures are what travels to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import c_oracle, encode_gt, ppn_oracle as O, ref_live, synth  # noqa: E402
from pytorch_pose_proposal_network_b200 import config as pcfg  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
STORE_INPUT_BELOW = 96 * 1024       # bytes


def geom_meta(g: O.Geometry) -> dict:
    return dict(K=g.K, E=g.E, inW=g.inW, inH=g.inH, W=g.W, H=g.H, sW=g.sW, sH=g.sH,
                graphs=[[list(a), list(b)] for a, b in g.graphs], det_thresh=g.det_thresh,
                nms_thresh=g.nms_thresh, min_kp=g.min_kp)


def pack_reference(humans, scores, K):
    n = len(humans)
    order = np.full((n, K), -1, np.int32)
    box = np.zeros((n, K, 4), np.float32)
    score = np.zeros((n, K), np.float32)
    for i, (hm, sc) in enumerate(zip(humans, scores)):
        assert list(hm.keys()) == list(sc.keys())
        for j, t in enumerate(hm.keys()):
            order[i, j] = t
            assert hm[t].dtype == np.float32 and hm[t].shape == (4,)
            box[i, t] = hm[t]
            score[i, t] = sc[t]
    return order, box, score


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def one_case(name, g: O.Geometry, out: np.ndarray, meta: dict):
    """out: [C,H,W]."""
    dt = ref_live.configure(g)
    assert synth.root_scores_distinct(out[None], g, g.det_thresh), f"{name}: duplicate root scores"
    # ---- the reference, end to end and stage by stage --------------------------------
    humans, scores = ref_live.reference_parse(out, g)
    K = g.K
    resp, conf, x, y, w, h, e = O.split_head(out, g)
    delta = resp * conf
    rx, ry = dt.restore_xy(x, y)
    rw, rh = dt.restore_size(w, h)
    ref_bbox0 = np.array([ry - rh / 2, rx - rw / 2, ry + rh / 2, rx + rw / 2]).transpose(1, 2, 3, 0)[0]
    cand = np.where(delta[0] > g.det_thresh)
    cand_cell = (cand[0] * g.W + cand[1]).astype(np.int32)
    ref_keep = dt.non_maximum_suppression(bbox=ref_bbox0[cand], thresh=g.nms_thresh, score=delta[0][cand])
    ref_amax = e.transpose(0, 3, 4, 1, 2).reshape(g.E, g.H, g.W, g.S).argmax(-1).astype(np.int32)
    order, box, score = pack_reference(humans, scores, K)

    # ---- numpy restatement must equal it ------------------------------------------------
    p = O.parse_image(out, g)
    assert np.array_equal(p.cand_cell, cand_cell), name
    assert np.array_equal(p.keep_idx, ref_keep), name
    assert np.array_equal(p.amax, ref_amax), name
    assert len(p.key_order) == len(humans), (name, len(p.key_order), len(humans))
    for i, ko in enumerate(p.key_order):
        assert ko == [int(t) for t in order[i] if t >= 0], name
    assert np.array_equal(bits(p.part_box), bits(box)), name
    assert np.array_equal(bits(p.part_score), bits(score)), name
    assert np.array_equal(bits(O.boxes(x, y, w, h, g)[0]), bits(ref_bbox0)), name
    hp, sp = O.parse_head_like_reference(out, g)
    o2, b2, s2 = pack_reference(hp, sp, K)
    assert np.array_equal(o2, order) and np.array_equal(bits(b2), bits(box)) and np.array_equal(bits(s2), bits(score))

    # ---- C restatement must equal it ----------------------------------------------------
    c = c_oracle.parse_batch(out[None], g)
    nc, nk, nh = (int(v) for v in c["counts"][0])
    assert nc == cand_cell.size and nk == ref_keep.size and nh == len(humans), (name, nc, nk, nh)
    assert np.array_equal(c["cand_cell"][0, :nc], cand_cell)
    assert np.array_equal(c["keep_idx"][0, :nk], ref_keep)
    assert np.array_equal(c["part_cell"][0, :nh], p.part_cell)
    assert np.array_equal(bits(c["part_box"][0, :nh]), bits(box))
    assert np.array_equal(bits(c["part_score"][0, :nh]), bits(score))
    assert np.array_equal(c_oracle.limb_argmax(out, g), ref_amax)

    fx = dict(
        meta=json.dumps({**meta, "geometry": geom_meta(g), "source": "reference datatest.get_humans_by_feature"}),
        input_sha256=synth.digest(out),
        cand_cell=cand_cell, ref_keep_idx=ref_keep.astype(np.int32),
        ref_key_order=order, ref_box=box, ref_score=score,
        part_cell=p.part_cell, root_cell=p.root_cell,
        amax_sha256=synth.digest(ref_amax.astype(np.uint16)),
    )
    if out.nbytes <= STORE_INPUT_BELOW:
        fx["input"] = out
        fx["amax"] = ref_amax.astype(np.uint16)
    np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **fx)
    print(f"{name:28s} cand={cand_cell.size:4d} keep={ref_keep.size:4d} humans={len(humans):4d} "
          f"parts/human={np.mean((p.part_cell >= 0).sum(1)) if len(humans) else 0:.1f}")


def tiny_geometry():
    # 5 parts, 4 limbs: 0->1->2, 1->3, 0->4 ; two track orders share the prefix 0->1
    graphs = (((0, 1), (1, 2)), ((0, 2), (1, 3)), ((3,), (4,)))
    return O.Geometry(K=5, E=4, inW=96, inH=96, W=6, H=6, sW=5, sH=5, graphs=graphs)


def nms_cases(dt):
    """Hand cases for the NMS function alone (datatest.py:134-160), answered by the reference."""
    rng = np.random.default_rng(7)
    cases = {}

    def add(name, box, thresh, score=None, limit=None):
        box = np.asarray(box, np.float32).reshape(-1, 4)
        sc = None if score is None else np.asarray(score, np.float32)
        with np.errstate(all="ignore"):
            keep = dt.non_maximum_suppression(box, thresh, score=sc, limit=limit)
        assert keep.dtype == np.int32
        for impl in (O.nms, c_oracle.nms):
            got = impl(box, thresh, score=sc, limit=limit)
            assert np.array_equal(got, keep), (name, impl.__module__, got, keep)
        cases[name] = dict(box=box, score=sc if sc is not None else np.zeros(0, np.float32),
                           has_score=sc is not None, thresh=float(thresh), limit=-1 if limit is None else limit, keep=keep)

    add("empty", np.zeros((0, 4)), 0.3, score=np.zeros(0))
    add("single", [[0, 0, 10, 10]], 0.3, score=[0.5])
    add("identical", [[0, 0, 10, 10]] * 3, 0.3, score=[0.5, 0.9, 0.7])
    add("disjoint", [[0, 0, 10, 10], [20, 20, 30, 30], [40, 0, 50, 10]], 0.3, score=[0.2, 0.3, 0.25])
    add("touching", [[0, 0, 10, 10], [10, 0, 20, 10], [0, 10, 10, 20]], 0.01, score=[0.9, 0.8, 0.7])
    add("zero_area_nan", [[5, 5, 5, 5], [5, 5, 5, 5], [0, 0, 10, 10]], 0.3, score=[0.9, 0.8, 0.7])
    add("inverted_box", [[10, 10, 0, 0], [0, 0, 10, 10], [1, 1, 9, 9]], 0.3, score=[0.9, 0.8, 0.7])
    add("nested", [[0, 0, 100, 100], [10, 10, 90, 90], [40, 40, 60, 60]], 0.5, score=[0.5, 0.6, 0.7])
    add("thresh_exact", [[0, 0, 10, 10], [0, 5, 10, 15]], 1.0 / 3.0, score=[0.9, 0.8])
    add("no_score_order", [[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30], [0, 0, 10, 10]], 0.3)
    for n, lim in ((40, None), (40, 3), (300, None), (1000, None), (1000, 17)):
        c = rng.random((n, 2), dtype=np.float32) * 300
        s = rng.random((n, 2), dtype=np.float32) * 80 + 4
        box = np.concatenate([c - s / 2, c + s / 2], axis=1)
        sc = rng.permutation(n).astype(np.float32) / n          # distinct scores
        add(f"random_n{n}_limit{lim}", box, 0.3, score=sc, limit=lim)
        add(f"random_n{n}_limit{lim}_noscore", box, 0.45, limit=lim)
    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            flat[f"{name}/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLDEN, "nms_cases.npz"), **flat)
    print(f"nms_cases: {len(cases)} cases")


def roundtrip_case(name, cfg, seed, n_people):
    """KAT-2: encode people as the reference's dataset does, parse the targets back."""
    g = O.Geometry.of(cfg)
    rng = np.random.default_rng(seed)
    edges = pcfg.EDGES if cfg.K == 18 else pcfg.EDGES_16
    gridW = g.inW // g.W
    people, taken = [], set()
    while len(people) < n_people:
        cx, cy = rng.uniform(0.2, 0.8, 2) * (g.inW, g.inH)
        cell = (int(cy // gridW), int(cx // gridW))
        if any(abs(cell[0] - t[0]) + abs(cell[1] - t[1]) < 3 for t in taken):
            continue
        taken.add(cell)
        bw, bh = rng.uniform(1.2, 2.0, 2) * gridW
        pts = {}
        for k in range(1, g.K):
            if rng.random() < 0.85:
                px, py = cx + rng.uniform(-1, 1) * 1.8 * gridW, cy + rng.uniform(-1, 1) * 1.8 * gridW
                pts[k] = (float(np.clip(px, 0, g.inW - 1)), float(np.clip(py, 0, g.inH - 1)))
        people.append({"box": (float(cx), float(cy), float(bw), float(bh)), "points": pts})
    out = encode_gt.encode_people(people, g, edges)
    # GT deltas are all exactly 1; give each person's root a distinct confidence so the
    # (reference-undefined) order of equal scores does not enter the fixture
    for i, person in enumerate(people):
        cx, cy = person["box"][:2]
        out[g.K + 0, int(cy / gridW), int(cx / gridW)] = np.float32(1.0 - 0.03 * i)
    one_case(name, g, out, dict(dist="roundtrip", seed=seed, n_people=n_people,
                                people=json.dumps(people)))
    return out


def pred_frame_cases(presets):
    """datatest.evaluation's prediction records (datatest.py:298-328) for two of the cases above."""
    cases = {}
    for name, g, dist, seed in (("tiny_U_s11", tiny_geometry(), "U", 11),
                                ("cfg2_U_s1", O.Geometry.of(presets["cfg2"]), "U", 1)):
        out = synth.make_head(g, dist, seed)[0]
        ref_live.configure(g)
        humans, scores = ref_live.reference_parse(out, g)
        if g.K != 18:
            # evaluation() loops over len(KEYPOINT_NAMES) of the reference's config (18): give it K names
            dt = ref_live.load()
            saved = dt.KEYPOINT_NAMES
            dt.KEYPOINT_NAMES = list(range(g.K))
        try:
            frames = ref_live.reference_pred_frames([name + ".jpg"], [humans], [scores])
        finally:
            if g.K != 18:
                dt.KEYPOINT_NAMES = saved
        want = O.canonical(frames[0])
        assert O.canonical(O.pred_frame(name + ".jpg", humans, scores, g.K)) == want, name
        p = O.parse_image(out, g)
        ho, so = O.humans_as_dicts(p)
        assert O.canonical(O.pred_frame(name + ".jpg", ho, so, g.K)) == want, name
        cases[name] = want
        print(f"pred_frame {name}: {len(frames[0]['annorect'])} persons")
    with open(os.path.join(GOLDEN, "pred_frames.json"), "w") as f:
        json.dump(cases, f)


def drawing_cases(presets):
    """The image draw_humans produces (datatest.py:162-232) for the humans of three 18-part cases: its sha256 goes
    into tests/golden/drawings.json.  Checked here: the primitives of oracle/skeleton.py, drawn by the package's
    drawing module, give the same pixels."""
    import hashlib
    from oracle import skeleton
    from PIL import Image
    from pytorch_pose_proposal_network_b200 import drawing
    cases = {}
    for name, preset, dist, seed in (("native_U_s0", "native", "U", 0), ("native_S_s3000", "native", "S", 3000),
                                     ("cfg4_U_s1", "cfg4", "U", 1)):
        g = O.Geometry.of(presets[preset])
        out = synth.make_head(g, dist, seed)[0]
        humans, _ = ref_live.reference_parse(out, g)
        p = O.parse_image(out, g)
        # the reference's draw_humans raises (PIL: "y1 must be greater than or equal to y0") on a root box less than
        # two pixels wide or high: such humans of the random tensors are left out, on both sides
        keep = [i for i, hm in enumerate(humans)
                if int(hm[0][3]) - int(hm[0][1]) >= 2 and int(hm[0][2]) - int(hm[0][0]) >= 2]
        humans = [humans[i] for i in keep]
        part_cell, part_box = p.part_cell[keep], p.part_box[keep]
        rect, kp, seg = skeleton.primitives(part_cell, part_box, pcfg.EDGES)
        rec = {"size": g.inW, "humans": len(humans), "kept": keep}
        for visbbox in (False, True):
            ref = ref_live.reference_draw(humans, g.inW, visbbox=visbbox)
            mine = np.asarray(drawing.draw_skeletons(Image.new("RGB", (g.inW, g.inW)), rect, kp, seg, pcfg.KEYPOINT_NAMES,
                                                     pcfg.EDGES, visbbox=visbbox, part_box=part_box))
            assert ref.shape == mine.shape and np.array_equal(ref, mine), (name, visbbox, int((ref != mine).sum()))
            rec["sha256_visbbox" if visbbox else "sha256"] = hashlib.sha256(ref.tobytes()).hexdigest()
        rec["primitives_sha256"] = hashlib.sha256(rect.tobytes() + kp.tobytes() + seg.tobytes()).hexdigest()
        cases[name] = rec
        print(f"drawing {name}: {len(humans)} humans")
    with open(os.path.join(GOLDEN, "drawings.json"), "w") as f:
        json.dump(cases, f, sort_keys=True)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    assert ref_live.available(), "run this where /root/reference exists"
    dt = ref_live.load()
    import config as refcfg                      # the reference's config.py
    assert refcfg.KEYPOINT_NAMES == pcfg.KEYPOINT_NAMES and refcfg.EDGES == pcfg.EDGES
    assert refcfg.DIRECTED_GRAPHS == pcfg.DIRECTED_GRAPHS and refcfg.EDGES_BY_NAME == pcfg.EDGES_BY_NAME
    assert refcfg.COLOR_MAP == pcfg.COLOR_MAP and list(refcfg.COLOR_MAP) == list(pcfg.COLOR_MAP)

    tg = tiny_geometry()
    for seed in (11, 12, 13):
        one_case(f"tiny_U_s{seed}", tg, synth.make_head(tg, "U", seed)[0], dict(dist="U", seed=seed))
    one_case("tiny_D_s14", tg, synth.make_head(tg, "D", 14)[0], dict(dist="D", seed=14))

    presets = {"cfg2": pcfg.PPNConfig.mpii16(), "cfg3": pcfg.PPNConfig.coco18(),
               "cfg4": pcfg.PPNConfig.highres(), "native": pcfg.PPNConfig.reference_native()}
    plan = [("cfg2", "U", 1), ("cfg2", "U", 2), ("cfg2", "R", 1000), ("cfg2", "S", 3000), ("cfg2", "D", 2001),
            ("cfg3", "U", 1), ("cfg3", "D", 2000), ("cfg3", "R", 1000),
            ("cfg4", "U", 1), ("cfg4", "D", 2002), ("cfg4", "S", 3000),
            ("native", "U", 0), ("native", "R", 1000), ("native", "S", 3000)]
    for cname, dist, seed in plan:
        g = O.Geometry.of(presets[cname])
        one_case(f"{cname}_{dist}_s{seed}", g, synth.make_head(g, dist, seed)[0],
                 dict(preset=cname, dist=dist, seed=seed))
    # keep root-only humans (the stale copy's min_num_keypoints=-1, test.py:159) — also valid for datatest's fn
    g = O.Geometry.of(presets["cfg2"].with_(min_num_keypoints=-1, detection_thresh=0.09))
    one_case("cfg2_U_s5_minkp-1_thr0.09", g, synth.make_head(g, "U", 5)[0],
             dict(preset="cfg2", dist="U", seed=5, min_num_keypoints=-1, detection_thresh=0.09))

    pred_frame_cases(presets)
    drawing_cases(presets)
    roundtrip_case("roundtrip_cfg2_s21", presets["cfg2"], 21, 3)
    roundtrip_case("roundtrip_native_s22", presets["native"], 22, 5)
    nms_cases(dt)


if __name__ == "__main__":
    main()
