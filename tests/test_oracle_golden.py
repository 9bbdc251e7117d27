"""The CPU restatements (numpy and C) against the reference's own outputs in tests/golden/."""
import numpy as np
import pytest

from oracle import c_oracle, ppn_oracle as O
from tests.golden_util import case_names, load_case, load_nms_cases
from oracle import synth


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", case_names())
def test_numpy_restatement_matches_reference(name):
    g, out, fx = load_case(name)
    p = O.parse_image(out, g)
    assert np.array_equal(p.cand_cell, fx["cand_cell"])
    assert np.array_equal(p.keep_idx, fx["ref_keep_idx"])
    assert synth.digest(p.amax.astype(np.uint16)) == str(fx["amax_sha256"])
    order = fx["ref_key_order"]
    assert len(p.key_order) == order.shape[0]
    for i, ko in enumerate(p.key_order):
        assert ko == [int(t) for t in order[i] if t >= 0]
    assert np.array_equal(bits(p.part_box), bits(fx["ref_box"]))
    assert np.array_equal(bits(p.part_score), bits(fx["ref_score"]))
    assert np.array_equal(p.part_cell, fx["part_cell"])


@pytest.mark.parametrize("name", case_names())
def test_reference_shaped_port_matches_reference(name):
    """The per-human-loop port that bench.py times as the CPU baseline."""
    g, out, fx = load_case(name)
    humans, scores = O.parse_head_like_reference(out, g)
    order = fx["ref_key_order"]
    assert len(humans) == order.shape[0]
    for i, (hm, sc) in enumerate(zip(humans, scores)):
        assert list(hm.keys()) == [int(t) for t in order[i] if t >= 0]
        for t in hm:
            assert np.array_equal(bits(hm[t]), bits(fx["ref_box"][i, t]))
            assert bits(np.float32(sc[t])) == bits(fx["ref_score"][i, t])


@pytest.mark.parametrize("name", case_names())
def test_c_restatement_matches_reference(name):
    g, out, fx = load_case(name)
    c = c_oracle.parse_batch(out[None], g, n_threads=1)
    nc, nk, nh = (int(v) for v in c["counts"][0])
    assert np.array_equal(c["cand_cell"][0, :nc], fx["cand_cell"])
    assert np.array_equal(c["keep_idx"][0, :nk], fx["ref_keep_idx"])
    assert nh == fx["ref_key_order"].shape[0]
    assert np.array_equal(c["part_cell"][0, :nh], fx["part_cell"])
    assert np.array_equal(bits(c["part_box"][0, :nh]), bits(fx["ref_box"]))
    assert np.array_equal(bits(c["part_score"][0, :nh]), bits(fx["ref_score"]))
    amax = c_oracle.limb_argmax(out, g)
    assert synth.digest(amax.astype(np.uint16)) == str(fx["amax_sha256"])


@pytest.mark.parametrize("impl", [O.nms, c_oracle.nms], ids=["numpy", "c"])
def test_nms_hand_cases(impl):
    for name, c in load_nms_cases().items():
        score = c["score"] if bool(c["has_score"]) else None
        limit = None if int(c["limit"]) < 0 else int(c["limit"])
        with np.errstate(all="ignore"):
            got = impl(c["box"], float(c["thresh"]), score=score, limit=limit)
        assert got.dtype == np.int32
        assert np.array_equal(got, c["keep"]), name


def test_c_batch_threads_agree():
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    g = O.Geometry.of(PPNConfig.mpii16())
    head = synth.make_head(g, "U", 99, B=6)
    a = c_oracle.parse_batch(head, g, n_threads=1)
    b = c_oracle.parse_batch(head, g, n_threads=4)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_on_demand_argmax_equals_dense():
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    g = O.Geometry.of(PPNConfig.mpii16())
    out = synth.make_head(g, "U", 3)[0]
    dense = c_oracle.limb_argmax(out, g)
    rng = np.random.default_rng(0)
    for _ in range(50):
        ei, c = int(rng.integers(g.E)), int(rng.integers(g.H * g.W))
        assert c_oracle.window_argmax(out, g, ei, c) == dense[ei].reshape(-1)[c]


def test_argmax_nan_and_tie_semantics():
    """numpy argmax: first maximum wins; a NaN is the maximum and the first NaN wins."""
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    g = O.Geometry.of(PPNConfig.mpii16(outsize=(4, 4), local_grid_size=(3, 3), insize=(128, 128)))
    out = np.zeros((g.C, g.H, g.W), np.float32)
    e = out[6 * g.K:].reshape(g.E, g.S, g.H * g.W)
    e[0, :, 0] = [1, 5, 5, 2, 5, 0, 0, 0, 0]            # tie -> 1
    e[0, :, 1] = [1, np.nan, 7, np.nan, 9, 0, 0, 0, 0]  # first NaN -> 1
    e[0, :, 2] = [-np.inf] * 9                          # all equal -> 0
    e[0, :, 3] = [-0.0, 0.0, -0.0, 0, 0, 0, 0, 0, 0]    # signed zeros equal -> 0
    e[0, :, 4] = [np.inf, np.nan, np.inf, 0, 0, 0, 0, 0, 0]  # NaN beats inf -> 1
    want = e.reshape(g.E, g.S, g.H, g.W).argmax(1)
    assert list(want[0].reshape(-1)[:5]) == [1, 1, 0, 0, 1]
    assert np.array_equal(c_oracle.limb_argmax(out, g), want)
    assert np.array_equal(O.limb_argmax(out[6 * g.K:].reshape(g.E, g.sH, g.sW, g.H, g.W)), want)


def test_pred_frames_match_reference():
    """datatest.evaluation's prediction records (datatest.py:298-328), captured from the reference."""
    import json
    import os
    from tests.golden_util import GOLDEN
    want = json.load(open(os.path.join(GOLDEN, "pred_frames.json")))
    for name, frame in want.items():
        g, out, _ = load_case(name)
        humans, scores = O.humans_as_dicts(O.parse_image(out, g))
        assert O.canonical(O.pred_frame(name + ".jpg", humans, scores, g.K)) == frame, name
