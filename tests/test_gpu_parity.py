"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden vectors.

Bar: bit-exact on candidate cells, NMS survivors, arg-max indices, human assignment AND on the
fp32 boxes / scores (the spec allows 1e-5 relative; the kernels reproduce numpy's roundings, so
the tests demand 0 ulp).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, ppn_oracle as O, synth
from tests.golden_util import case_names, load_case, load_nms_cases

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def cfg_of(g: O.Geometry):
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    return PPNConfig(K=g.K, E=g.E, insize=(g.inW, g.inH), outsize=(g.W, g.H), local_grid_size=(g.sW, g.sH),
                     directed_graphs=g.graphs, detection_thresh=g.det_thresh, nms_thresh=g.nms_thresh,
                     min_num_keypoints=g.min_kp)


def parser_for(g):
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    return PoseParser(cfg_of(g))


def assert_packed_equals_oracle(packed, ref, B):
    """packed: PackedHumans.numpy(); ref: c_oracle.parse_batch dict."""
    assert np.array_equal(packed["count"], ref["counts"][:, 2])
    for b in range(B):
        n = int(ref["counts"][b, 2])
        assert np.array_equal(packed["root_cell"][b, :n], ref["root_cell"][b, :n]), b
        assert np.array_equal(packed["part_cell"][b, :n], ref["part_cell"][b, :n]), b
        assert np.array_equal(bits(packed["part_score"][b, :n]), bits(ref["part_score"][b, :n])), b
        assert np.array_equal(bits(packed["part_box"][b, :n]), bits(ref["part_box"][b, :n])), b


# ------------------------------------------------------------------------------------------
# golden vectors: outputs of the reference's own functions
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", case_names())
def test_golden_stages_and_humans(name):
    g, out, fx = load_case(name)
    parser = parser_for(g)
    head = torch.from_numpy(out[None]).cuda()

    amax = parser.limb_argmax(head)
    assert synth.digest(amax[0].cpu().numpy()) == str(fx["amax_sha256"])

    cell, score, box, count = parser.decode_candidates(head)
    n = int(count[0, 0])
    assert np.array_equal(cell[0, 0, :n].cpu().numpy(), fx["cand_cell"])

    keep, kcount = parser.nms(box[:, 0], score[:, 0], count[:, 0].contiguous(), g.nms_thresh)
    m = int(kcount[0])
    assert np.array_equal(keep[0, :m].cpu().numpy(), fx["ref_keep_idx"])

    staged = parser.tree_parse(head, amax, cell, keep.reshape(1, 1, -1), kcount.reshape(1, 1))
    fused = parser.parse(head)
    torch.cuda.synchronize()
    order = fx["ref_key_order"]
    for packed in (staged, fused):
        a = packed.numpy()
        nh = int(a["count"][0])
        assert nh == order.shape[0]
        assert np.array_equal(a["part_cell"][0, :nh], fx["part_cell"])
        assert np.array_equal(a["root_cell"][0, :nh], fx["root_cell"])
        assert np.array_equal(bits(a["part_box"][0, :nh]), bits(fx["ref_box"]))
        assert np.array_equal(bits(a["part_score"][0, :nh]), bits(fx["ref_score"]))
        humans, scores = packed.humans(0)
        for i, (hm, sc) in enumerate(zip(humans, scores)):
            want = [int(t) for t in order[i] if t >= 0]
            assert list(hm.keys()) == want and list(sc.keys()) == want
            for t in want:
                assert hm[t].dtype == np.float32 and hm[t].shape == (4,)
                assert isinstance(sc[t], np.float32)


def test_golden_dropin_signature_native():
    """datatest.get_humans_by_feature with the reference's own argument convention (numpy, squeezed)."""
    from pytorch_pose_proposal_network_b200 import datatest as dt
    g, out, fx = load_case("native_U_s0")
    resp, conf, x, y, w, h, e = O.split_head(out, g)
    humans, scores = dt.get_humans_by_feature(resp * conf, x, y, w, h, e, detection_thresh=0.15)
    order = fx["ref_key_order"]
    assert len(humans) == order.shape[0]
    for i, (hm, sc) in enumerate(zip(humans, scores)):
        want = [int(t) for t in order[i] if t >= 0]
        assert list(hm.keys()) == want
        for t in want:
            assert np.array_equal(bits(hm[t]), bits(fx["ref_box"][i, t]))
            assert bits(sc[t]) == bits(fx["ref_score"][i, t])


def test_dropin_test_py_variant():
    """test.py's stale copy: thresholds 0.09 / NMS 0.5 / min_num_keypoints=-1, resp*conf inside,
    geometry from `model`, humans only (test.py:159-221) — same kernels, other parameters."""
    from types import SimpleNamespace
    from pytorch_pose_proposal_network_b200 import variant_test_py as tv
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    cfg = PPNConfig.reference_native(detection_thresh=0.09, nms_thresh=0.5, min_num_keypoints=-1, swap_window_offsets=True)
    g = O.Geometry.of(cfg)
    out = synth.make_head(g, "U", seed=8, B=1)[0]
    want_h, want_s = O.humans_as_dicts(O.parse_image(out, g))
    resp, conf, x, y, w, h, e = O.split_head(out, g)
    model = SimpleNamespace(insize=(384, 384), outsize=(24, 24), local_grid_size=(21, 21))
    humans = tv.get_humans_by_feature(model, resp, conf, x, y, w, h, e)
    assert len(humans) == len(want_h) and any(len(hm) == 1 for hm in humans)      # root-only humans are kept
    for hm, wh in zip(humans, want_h):
        assert list(hm.keys()) == list(wh.keys())
        for t in hm:
            assert np.array_equal(bits(hm[t]), bits(wh[t]))


def test_dropin_rt_test_inference():
    """rt_test.inference with a stand-in model: same preprocessing, GPU parse instead of 7 copies + numpy."""
    from pytorch_pose_proposal_network_b200 import rt_test
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    cfg = PPNConfig.reference_native()
    g = O.Geometry.of(cfg)
    fixed = torch.from_numpy(synth.make_head(g, "U", seed=9, B=1)).cuda()

    class Head(torch.nn.Module):
        insize, outsize, local_grid_size = (384, 384), (24, 24), (21, 21)
        def forward(self, x):
            assert x.shape == (1, 3, 384, 384) and x.dtype == torch.float32
            return fixed
    image = (np.random.default_rng(0).random((384, 384, 3)) * 255).astype(np.uint8)
    humans, scores = rt_test.inference(image, Head(), (24, 24), (21, 21), return_humans=True)
    want_h, want_s = O.humans_as_dicts(O.parse_image(fixed[0].cpu().numpy(), g))
    assert len(humans) == len(want_h) > 0
    for hm, wh, sc, ws in zip(humans, want_h, scores, want_s):
        assert list(hm.keys()) == list(wh.keys())
        for t in hm:
            assert np.array_equal(bits(hm[t]), bits(wh[t])) and bits(sc[t]) == bits(ws[t])
    drawn = rt_test.inference(image, Head(), (24, 24), (21, 21), draw=lambda **kw: (kw["pil_image"].size, len(kw["humans"])))
    assert drawn == ((384, 384), len(want_h))


def test_part_centres_and_evaluation_frames():
    """ppn_part_centres + evaluation.pred_frame against the frames the reference's evaluation() builds."""
    import json
    import os
    from tests.golden_util import GOLDEN
    from pytorch_pose_proposal_network_b200 import evaluation
    want = json.load(open(os.path.join(GOLDEN, "pred_frames.json")))
    for name, frame in want.items():
        g, out, _ = load_case(name)
        parser = parser_for(g)
        packed = parser.parse(torch.from_numpy(out[None]).cuda())
        centres = parser.part_centres(packed)
        got = evaluation.pred_frames([name + ".jpg"], packed, centres)[0]
        assert O.canonical(got) == frame, name
        # unused slots and absent parts are zero
        c = centres.cpu().numpy()[0]
        a = packed.numpy()
        n = int(a["count"][0])
        assert not c[n:].any() and not c[:n][a["part_cell"][0, :n] < 0].any()


def test_dropin_restore_functions():
    from pytorch_pose_proposal_network_b200 import datatest as dt
    g, out, _ = load_case("native_U_s0")
    _, _, x, y, w, h, _ = O.split_head(out, g)
    rx, ry = dt.restore_xy(x, y)
    rw, rh = dt.restore_size(w, h)
    ex, ey = O.restore_xy(x, y, g)
    ew, eh = O.restore_size(w, h, g)
    for got, want in ((rx, ex), (ry, ey), (rw, ew), (rh, eh)):
        assert got.dtype == np.float32 and np.array_equal(bits(got), bits(want))


def test_dropin_nms_hand_cases():
    from pytorch_pose_proposal_network_b200 import datatest as dt
    for name, c in load_nms_cases().items():
        score = c["score"] if bool(c["has_score"]) else None
        limit = None if int(c["limit"]) < 0 else int(c["limit"])
        got = dt.non_maximum_suppression(c["box"], float(c["thresh"]), score=score, limit=limit)
        assert got.dtype == np.int32
        assert np.array_equal(got, c["keep"]), name


def test_nms_long_list_uses_global_kernel():
    """> 1024 boxes: the no-bitmask kernel; checked against the C restatement."""
    from pytorch_pose_proposal_network_b200 import datatest as dt
    rng = np.random.default_rng(5)
    n = 3000
    c = rng.random((n, 2), dtype=np.float32) * 600
    s = rng.random((n, 2), dtype=np.float32) * 60 + 4
    box = np.concatenate([c - s / 2, c + s / 2], axis=1)
    score = rng.permutation(n).astype(np.float32)
    for sc, lim in ((score, None), (None, None), (score, 40)):
        got = dt.non_maximum_suppression(box, 0.3, score=sc, limit=lim)
        assert np.array_equal(got, c_oracle.nms(box, 0.3, score=sc, limit=lim))


@pytest.mark.parametrize("blockwise", [0, 1], ids=["wavefront", "blockwise"])
def test_nms_both_phase3_implementations(blockwise):
    """The warp-wavefront NMS (default) and the block-by-block one (`nms.blockwise`) against the C restatement:
    list lengths around the 32-box words and the 16-warp wrap, crowded and sparse lists, limits, NaN boxes."""
    from pytorch_pose_proposal_network_b200 import _lib, datatest as dt
    rng = np.random.default_rng(77)
    _lib.tune(nms_blockwise=blockwise)
    try:
        for n in (1, 2, 31, 32, 33, 64, 65, 200, 511, 512, 513, 576, 1000, 1024):
            for spread, size in ((40.0, 30.0), (400.0, 30.0), (4000.0, 20.0)):
                c = rng.random((n, 2), dtype=np.float32) * np.float32(spread)
                sz = rng.random((n, 2), dtype=np.float32) * np.float32(size) + 2
                box = np.concatenate([c - sz / 2, c + sz / 2], axis=1).astype(np.float32)
                score = rng.permutation(n).astype(np.float32)
                for lim in (None, 1, 7, 40):
                    got = dt.non_maximum_suppression(box, 0.3, score=score, limit=lim)
                    assert np.array_equal(got, c_oracle.nms(box, 0.3, score=score, limit=lim)), (n, spread, lim)
                nb = box.copy()
                nb[rng.integers(0, n, max(1, n // 17)), rng.integers(0, 4, max(1, n // 17))] = np.nan
                got = dt.non_maximum_suppression(nb, 0.3, score=score)
                assert np.array_equal(got, c_oracle.nms(nb, 0.3, score=score)), (n, spread, "nan")
                got = dt.non_maximum_suppression(box, 0.0)            # threshold 0: no early-out, everything overlapping goes
                assert np.array_equal(got, c_oracle.nms(box, 0.0)), (n, spread, "thr0")
    finally:
        _lib.tune(nms_blockwise=0)


# ------------------------------------------------------------------------------------------
# batches against the C restatement, every distribution and preset
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("preset,dist,B", [("cfg2", "U", 48), ("cfg2", "R", 16), ("cfg2", "D", 16), ("cfg2", "S", 16),
                                           ("cfg3", "D", 12), ("cfg3", "U", 12), ("cfg4", "U", 6), ("cfg4", "D", 4),
                                           ("native", "U", 3)])
def test_batch_matches_c_oracle(preset, dist, B):
    from pytorch_pose_proposal_network_b200.config import PRESETS
    cfg = PRESETS[preset]()
    g = O.Geometry.of(cfg)
    from tests.gpu_inputs import device_head
    dev, head = device_head(g, dist, seed=100 + B, B=B)       # generated on the device; the oracle gets the host copy
    assert synth.root_scores_distinct(head, g)
    ref = c_oracle.parse_batch(head, g, n_threads=8)
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    parser = PoseParser(cfg)
    packed = parser.parse(dev)
    assert_packed_equals_oracle(packed.numpy(), ref, B)
    # the host-memory entry (chunked, overlapped copies) must give the same bytes
    host = parser.parse_host(torch.from_numpy(head).pin_memory())
    assert_packed_equals_oracle(host.numpy(), ref, B)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("preset,dist,B,grid", [("cfg2", "U", 24, None), ("cfg2", "D", 8, None), ("cfg3", "D", 6, None),
                                               ("cfg4", "U", 4, None), ("cfg2", "U", 5, (13, 13)), ("cfg2", "R", 5, (10, 6))])
def test_sixteen_bit_head(preset, dist, B, grid, dtype):
    """fp16 / bf16 head tensors: every element is widened exactly, so the result must equal the fp32
    restatement run on head.float() — including ties, which 16-bit values produce in abundance
    (first arg-max wins; equal root scores: larger cell first, in oracle and kernel alike)."""
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS[preset]()
    if grid:
        cfg = cfg.with_(outsize=grid, insize=(grid[0] * 16, grid[1] * 16))
    g = O.Geometry.of(cfg)
    from tests.gpu_inputs import device_head
    dev16, up = device_head(g, dist, seed=77, B=B, dtype=dtype, distinct=False)           # rounded to 16 bits on the device; `up` = widened host copy
    head16 = dev16.cpu()
    ref = c_oracle.parse_batch(up, g, n_threads=8)
    parser = PoseParser(cfg)
    dev = head16.cuda()
    want_amax = np.stack([c_oracle.limb_argmax(img, g) for img in up]).astype(np.uint16)
    assert np.array_equal(parser.limb_argmax(dev).cpu().numpy(), want_amax)
    assert_packed_equals_oracle(parser.parse(dev).numpy(), ref, B)
    assert_packed_equals_oracle(parser.parse(dev, input_complete=True).numpy(), ref, B)
    cell, score, box, count = parser.decode_candidates(dev)
    for b in range(B):
        n = int(count[b, 0])
        assert n == ref["counts"][b, 0] and np.array_equal(cell[b, 0, :n].cpu().numpy(), ref["cand_cell"][b, :n])
    host = parser.parse_host(head16.pin_memory())
    assert_packed_equals_oracle(host.numpy(), ref, B)


TUNE_DEFAULTS = dict(argmax_variant=0, argmax_stage_bytes=0, argmax_stages=4, argmax_threads=320,
                     argmax_ctas_per_sm=1, argmax_split=-1, argmax_dynamic=1, argmax_tail_opt=0)


@pytest.mark.parametrize("variant,stage_bytes,stages,threads,ctas,split", [
    (0, 32768, 5, 320, 1, -1), (0, 4096, 3, 96, 2, 0), (0, 65536, 2, 640, 1, 0), (0, 16384, 8, 256, 2, 0),
    (0, 32768, 4, 320, 1, 1), (0, 8192, 3, 160, 2, 1), (0, 65536, 3, 992, 1, 1), (0, 2048, 6, 64, 1, 1),
    (1, 32768, 5, 320, 1, -1), (1, 32768, 5, 64, 1, 1)])
@pytest.mark.parametrize("preset", ["cfg2", "cfg4", "native"])
def test_limb_argmax_every_tuning(preset, variant, stage_bytes, stages, threads, ctas, split):
    """The arg-max kernel under every ring shape / thread mapping, against the C restatement."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS[preset]()
    g = O.Geometry.of(cfg)
    B = 3 if preset == "native" else 7          # 7 x 15 = 105 matrices: a ragged last item in split-matrix mode
    head = synth.make_head(g, "U", seed=7, B=B)
    want = np.stack([c_oracle.limb_argmax(img, g) for img in head]).astype(np.uint16)
    _lib.tune(argmax_variant=variant, argmax_stage_bytes=stage_bytes, argmax_stages=stages, argmax_threads=threads,
              argmax_ctas_per_sm=ctas, argmax_split=split)
    try:
        got = PoseParser(cfg).limb_argmax(torch.from_numpy(head).cuda()).cpu().numpy()
    finally:
        _lib.tune(**TUNE_DEFAULTS)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("dynamic,tail_opt", [(1, 0), (1, 2), (1, 4), (0, 4), (1, 8)])
def test_limb_argmax_shrinking_tail(dtype, dynamic, tail_opt):
    """A batch of several waves of work items: the last waves are handed out as smaller items holding more rows of
    fewer matrices (plan_tail).  Every matrix must still be reduced exactly once, whatever the item sizes."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS["cfg2"]()
    g = O.Geometry.of(cfg)
    B = 301                                                       # 4515 matrices: 3-4 waves of 8-matrix items on 148 SMs, ragged end
    head = torch.from_numpy(synth.make_head(g, "U", seed=31, B=B)).to(dtype)
    up = head.float().numpy()
    want = np.stack([c_oracle.limb_argmax(img, g) for img in up]).astype(np.uint16)
    _lib.tune(argmax_dynamic=dynamic, argmax_tail_opt=tail_opt)
    try:
        parser = PoseParser(cfg)
        dev = head.cuda()
        for rep in range(2):
            assert np.array_equal(parser.limb_argmax(dev).cpu().numpy(), want), rep
        ref = c_oracle.parse_batch(up, g, n_threads=8)
        assert_packed_equals_oracle(parser.parse(dev, input_complete=True).numpy(), ref, B)
    finally:
        _lib.tune(**TUNE_DEFAULTS)


@pytest.mark.parametrize("preset", ["cfg2", "cfg4"])
@pytest.mark.parametrize("dynamic,tail_opt,ctas", [(0, 0, 1), (0, 2, 2), (1, 0, 1), (1, 4, 2)])
def test_limb_argmax_work_distribution(preset, dynamic, tail_opt, ctas):
    """Static round-robin vs ticket scheduling, repeated launches (the counter must reset itself)."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS[preset]()
    g = O.Geometry.of(cfg)
    head = synth.make_head(g, "U", seed=21, B=40 if preset == "cfg2" else 12)
    want = np.stack([c_oracle.limb_argmax(img, g) for img in head]).astype(np.uint16)
    dev = torch.from_numpy(head).cuda()
    _lib.tune(argmax_dynamic=dynamic, argmax_tail_opt=tail_opt, argmax_ctas_per_sm=ctas)
    try:
        parser = PoseParser(cfg)
        side = torch.cuda.Stream()
        for rep in range(4):
            got = parser.limb_argmax(dev)
            with torch.cuda.stream(side):                  # a second stream gets its own counter pair
                got2 = parser.limb_argmax(dev)
            torch.cuda.synchronize()
            assert np.array_equal(got.cpu().numpy(), want), rep
            assert np.array_equal(got2.cpu().numpy(), want), rep
    finally:
        _lib.tune(**TUNE_DEFAULTS)


@pytest.fixture(params=[1, 0], ids=["fused", "three_kernels"])
def fused(request):
    """Run a whole-path test with the default two-kernel call (arg-max + fused parse) and with the
    three-kernel chain (ppn_tune parse.fused = 0)."""
    from pytorch_pose_proposal_network_b200 import _lib
    _lib.tune(parse_fused=request.param)
    yield request.param
    _lib.tune(parse_fused=-1)


@pytest.mark.parametrize("preset,dist,B", [("cfg2", "U", 40), ("cfg3", "D", 24), ("cfg4", "U", 10)])
def test_overlapped_consecutive_calls(preset, dist, B, fused):
    """PPN_FLAG_INPUT_COMPLETE: back-to-back calls overlap (call i's tree parse under call i+1's
    arg-max, alternating workspace sets).  Every call's result must still be exact, with inputs and
    outputs changing from call to call and the modes interleaved."""
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS[preset]()
    g = O.Geometry.of(cfg)
    n_in = 5
    heads = [synth.make_head(g, dist, seed=500 + i, B=B) for i in range(n_in)]
    refs = [c_oracle.parse_batch(h, g, n_threads=8) for h in heads]
    devs = [torch.from_numpy(h).cuda() for h in heads]
    parser = PoseParser(cfg)
    n_calls = 23
    outs = [parser.alloc_output(B) for _ in range(n_calls)]
    torch.cuda.synchronize()
    for i in range(n_calls):
        parser.parse(devs[i % n_in], out=outs[i], input_complete=(i % 7 != 3))     # mostly overlapped, some not
    torch.cuda.synchronize()
    for i in range(n_calls):
        assert_packed_equals_oracle(outs[i].numpy(), refs[i % n_in], B)
    # same output buffer reused by overlapping calls: the last writer must win cleanly
    for i in range(6):
        last = parser.parse(devs[i % n_in], out=outs[0], input_complete=True)
    torch.cuda.synchronize()
    assert_packed_equals_oracle(last.numpy(), refs[5 % n_in], B)


@pytest.mark.parametrize("persist", [0, 1, 2])
@pytest.mark.parametrize("preset,dist,B", [("cfg2", "U", 400), ("cfg3", "D", 310)])
def test_three_kernel_chain_persistent_grids(preset, dist, B, persist):
    """The three-kernel chain with more images than resident CTAs: decode+NMS and the tree parse are persistent grids
    whose CTAs stride over the images (parse.persist CTAs per SM; 0 = one CTA per image), overlapped calls guarded by
    the published sequence number.  Every call's result must be exact."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS[preset]()
    g = O.Geometry.of(cfg)
    heads = [synth.make_head(g, dist, seed=700 + i, B=B) for i in range(3)]
    refs = [c_oracle.parse_batch(h, g, n_threads=8) for h in heads]
    devs = [torch.from_numpy(h).cuda() for h in heads]
    _lib.tune(parse_fused=0, parse_persist=persist)
    try:
        parser = PoseParser(cfg)
        assert parser.parse_plan(B)["launches"] == 3
        n_calls = 14
        outs = [parser.alloc_output(B) for _ in range(n_calls)]
        torch.cuda.synchronize()
        for i in range(n_calls):
            parser.parse(devs[i % 3], out=outs[i], input_complete=(i % 5 != 2))
        torch.cuda.synchronize()
        for i in range(n_calls):
            assert_packed_equals_oracle(outs[i].numpy(), refs[i % 3], B)
    finally:
        _lib.tune(parse_fused=-1, parse_persist=1)


def test_overlapped_chain_stress():
    """The overlapped chain under everything that can interleave with it on one stream: full-size
    batches (so that several calls' kernels really are in flight), changing batch sizes, calls
    without the flag, the three-kernel path (two NMS parts -> not fusable), stand-alone arg-max
    launches, a second parser with its own workspace, and foreign kernels in between.  Every one of
    result must be exact."""
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS["cfg2"]()
    g = O.Geometry.of(cfg)
    sizes = [512, 64, 512, 7, 300]
    heads = [synth.make_head(g, "U", seed=900 + i, B=n) for i, n in enumerate(sizes)]
    refs = [c_oracle.parse_batch(h, g, n_threads=8) for h in heads]
    devs = [torch.from_numpy(h).cuda() for h in heads]
    pa, pb, p2 = PoseParser(cfg), PoseParser(cfg), PoseParser(cfg, n_nms_parts=2)
    n_calls = 150
    outs, which = [], []
    scratch = torch.zeros(1 << 20, device="cuda")
    torch.cuda.synchronize()
    for i in range(n_calls):
        k = (i * 7 + i // 11) % len(sizes)
        parser = pb if i % 13 == 5 else pa
        outs.append(parser.parse(devs[k], out=parser.alloc_output(sizes[k]), input_complete=(i % 17 != 9)))
        which.append(k)
        if i % 29 == 3:
            pa.limb_argmax(devs[k])                           # stand-alone launch on the same stream
        if i % 31 == 4:
            scratch.add_(1.0)                                 # a foreign kernel between two calls
        if i % 37 == 6:
            outs.append(p2.parse(devs[k], out=p2.alloc_output(sizes[k]), input_complete=True))   # three-kernel path
            which.append(k)
    torch.cuda.synchronize()
    for o, k in zip(outs, which):
        assert_packed_equals_oracle(o.numpy(), refs[k], sizes[k])


def test_overlapped_chain_every_result():
    """120 overlapped full-size calls into 120 distinct outputs, all verified."""
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS["cfg2"]()
    g = O.Geometry.of(cfg)
    B = 256
    heads = [synth.make_head(g, "U", seed=950 + i, B=B) for i in range(3)]
    refs = [c_oracle.parse_batch(h, g, n_threads=8) for h in heads]
    devs = [torch.from_numpy(h).cuda() for h in heads]
    parser = PoseParser(cfg)
    outs = [parser.alloc_output(B) for _ in range(120)]
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        parser.parse(devs[i % 3], out=o, input_complete=True)
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        assert_packed_equals_oracle(o.numpy(), refs[i % 3], B)



def test_captured_parse_replays_on_refilled_buffer():
    """PoseParser.capture: one graph launch per frame on a buffer the caller refills (the rt_test.py loop)."""
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS["native"]()
    g = O.Geometry.of(cfg)
    parser = PoseParser(cfg)
    head = torch.zeros(1, cfg.C, cfg.H, cfg.W, device="cuda")
    cap = parser.capture(head)
    for seed in (5, 6, 7):
        frame = synth.make_head(g, "S" if seed == 6 else "U", seed=seed, B=1)
        head.copy_(torch.from_numpy(frame))
        packed = cap.replay()
        torch.cuda.synchronize()
        assert_packed_equals_oracle(packed.numpy(), c_oracle.parse_batch(frame, g), 1)


@pytest.mark.parametrize("overlap", [1, 2])
def test_cuda_graph_capture(overlap, fused):
    """The whole path is capturable: three parses (PDL chain or side-stream fork/join) recorded into
    one CUDA graph and replayed on new data."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PPNConfig.mpii16()
    g = O.Geometry.of(cfg)
    B = 12
    heads = [synth.make_head(g, "U", seed=700 + i, B=B) for i in range(2)]
    refs = [c_oracle.parse_batch(h, g, n_threads=8) for h in heads]
    static_in = torch.from_numpy(heads[0]).cuda()
    _lib.tune(parse_overlap=overlap)
    try:
        parser = PoseParser(cfg)
        outs = [parser.alloc_output(B) for _ in range(3)]
        parser.parse(static_in, out=outs[0])                 # first call outside capture (one-time set-up)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for o in outs:
                parser.parse(static_in, out=o, input_complete=True)
        for which in (1, 0, 1):
            static_in.copy_(torch.from_numpy(heads[which]).cuda())
            graph.replay()
            torch.cuda.synchronize()
            for o in outs:
                assert_packed_equals_oracle(o.numpy(), refs[which], B)
    finally:
        _lib.tune(parse_overlap=2)


def test_every_launch_ordering_gives_same_result(fused):
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PPNConfig.mpii16()
    g = O.Geometry.of(cfg)
    head = synth.make_head(g, "U", seed=9, B=20)
    ref = c_oracle.parse_batch(head, g, n_threads=8)
    dev = torch.from_numpy(head).cuda()
    for overlap in (0, 1, 2):
        _lib.tune(parse_overlap=overlap)
        try:
            parser = PoseParser(cfg)
            for _ in range(3):                     # back-to-back calls reuse the workspace
                packed = parser.parse(dev)
            assert_packed_equals_oracle(packed.numpy(), ref, 20)
        finally:
            _lib.tune(parse_overlap=2)


@pytest.mark.parametrize("W,H,sW,sH", [(13, 13, 9, 9), (5, 7, 3, 5), (10, 6, 7, 7), (31, 33, 3, 3)])
def test_odd_shapes_generic_paths(W, H, sW, sH):
    """Grids whose cell count is not a multiple of 4 (scalar arg-max kernel), non-square grids and
    windows (row/column half-window convention of datatest.py:115-116)."""
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PPNConfig.mpii16(insize=(W * 16, H * 16), outsize=(W, H), local_grid_size=(sW, sH))
    g = O.Geometry.of(cfg)
    head = synth.make_head(g, "U", seed=W * 100 + H, B=5)
    ref = c_oracle.parse_batch(head, g)
    parser = PoseParser(cfg)
    dev = torch.from_numpy(head).cuda()
    want = np.stack([c_oracle.limb_argmax(img, g) for img in head]).astype(np.uint16)
    assert np.array_equal(parser.limb_argmax(dev).cpu().numpy(), want)
    assert_packed_equals_oracle(parser.parse(dev).numpy(), ref, 5)


def test_argmax_nan_tie_signed_zero():
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PPNConfig.mpii16(outsize=(4, 4), local_grid_size=(3, 3), insize=(128, 128))
    g = O.Geometry.of(cfg)
    out = np.zeros((1, g.C, g.H, g.W), np.float32)
    e = out[0, 6 * g.K:].reshape(g.E, g.S, g.H * g.W)
    rng = np.random.default_rng(0)
    e[:] = rng.integers(0, 3, e.shape).astype(np.float32)            # plenty of exact ties
    e[0, :, 0] = [1, 5, 5, 2, 5, 0, 0, 0, 0]
    e[0, :, 1] = [1, np.nan, 7, np.nan, 9, 0, 0, 0, 0]
    e[0, :, 2] = [-np.inf] * 9
    e[0, :, 3] = [-0.0, 0.0, -0.0, 0, 0, 0, 0, 0, 0]
    e[0, :, 4] = [np.inf, np.nan, np.inf, 0, 0, 0, 0, 0, 0]
    e[1, :, 5] = [np.nan] * 9
    want = e.reshape(g.E, g.S, g.H, g.W).argmax(1).astype(np.uint16)
    from pytorch_pose_proposal_network_b200 import _lib
    for split in (0, 1):                  # row-split (merge path) and matrix-split mappings
        _lib.tune(argmax_split=split)
        try:
            got = PoseParser(cfg).limb_argmax(torch.from_numpy(out).cuda()).cpu().numpy()[0]
        finally:
            _lib.tune(argmax_split=-1)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("grid,window", [((4, 4), (3, 3)), ((8, 8), (5, 5)), ((12, 12), (9, 9))])
def test_argmax_sixteen_bit_special_values(dtype, grid, window):
    """The packed 16-bit comparison (HSET2 / HMNMX2 on raw halves) against numpy's arg-max of the
    up-cast tensor: NaNs (first one wins, either sign, any payload), infinities, signed zeros,
    denormals and long runs of ties, in the ring kernel (H*W % 8 == 0) over several ring shapes."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PPNConfig.mpii16(outsize=grid, local_grid_size=window, insize=(grid[0] * 32, grid[1] * 32))
    g = O.Geometry.of(cfg)
    B = 3
    rng = np.random.default_rng(5)
    # raw 16-bit patterns: every class of value appears, denormals and NaN payloads included
    bits = rng.integers(0, 1 << 16, (B, g.C, g.H, g.W), dtype=np.uint16)
    head16 = torch.from_numpy(bits.view(np.int16)).view(dtype).clone()
    e = head16[:, 6 * g.K:].view(B, g.E, g.S, g.H * g.W)          # a view: the writes below land in head16
    small = torch.from_numpy(rng.integers(0, 3, tuple(e.shape[1:])).astype(np.float32)).to(dtype)
    e[1] = small                                                          # image 1: exact ties everywhere
    tiny = torch.tensor([0.0, -0.0, 6e-8, -6e-8, 1e-40, -1e-40], dtype=torch.float32).to(dtype)
    e[2] = tiny[torch.from_numpy(rng.integers(0, len(tiny), tuple(e.shape[1:])))]   # zeros and denormals only
    e[2, 0, :, 0] = float("-inf")
    e[2, 0, :, 1] = float("nan")
    e[2, 0, 2:, 2] = float("inf")
    e[2, 1, g.S - 1, 3] = float("nan")                                   # NaN in the very last row
    up = head16.float().numpy()
    want = up[:, 6 * g.K:].reshape(B, g.E, g.S, g.H, g.W).argmax(2).astype(np.uint16)
    dev = head16.cuda()
    for stage_bytes, stages, threads in [(49152, 4, 480), (2048, 3, 64), (8192, 5, 992)]:
        _lib.tune(argmax16_stage_bytes=stage_bytes, argmax_stages=stages, argmax16_threads=threads)
        try:
            got = PoseParser(cfg).limb_argmax(dev).cpu().numpy()
        finally:
            _lib.tune(argmax16_stage_bytes=49152, argmax_stages=4, argmax16_threads=320)
        assert np.array_equal(got, want), (stage_bytes, stages, threads)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("preset,B,cluster", [("cfg2", 1, -1), ("cfg2", 3, -1), ("native", 1, -1), ("native", 2, 4), ("cfg4", 1, 8),
                                             ("cfg2", 2, 2), ("cfg3", 5, 8)])
@pytest.mark.parametrize("ring", [1, 0])
def test_limb_argmax_cluster_kernel_tiny_batches(preset, B, cluster, dtype, ring):
    """Tiny batches take the thread-block-cluster kernel (C CTAs per matrix, partials merged through
    distributed shared memory): same arg-max map as numpy's, ties across CTAs and NaNs included, and the
    whole path on top of it.  ring = 1: the rows stream through a ring of bulk copies; 0: 128-bit loads."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS[preset]()
    g = O.Geometry.of(cfg)
    head = torch.from_numpy(synth.make_head(g, "U", seed=70 + B, B=B)).to(dtype)
    e = head[:, 6 * g.K:].view(B, g.E, g.S, g.H * g.W)
    e[0, 0, :, 0] = 0.5                                  # a whole column of ties: row 0 must win across all CTAs
    e[0, 0, g.S // 2:, 1] = 2.0                          # tie that starts in a later CTA's rows
    e[0, 1, g.S - 1, 2] = float("nan")                   # NaN in the last CTA's rows
    e[0, 1, 1, 3] = float("nan"); e[0, 1, g.S - 2, 3] = float("nan")     # two NaNs: the first one wins
    up = head.float().numpy()
    want = up[:, 6 * g.K:].reshape(B, g.E, g.S, g.H, g.W).argmax(2).astype(np.uint16)
    _lib.tune(argmax_cluster=cluster, argmax_cluster_ring=ring)
    try:
        parser = PoseParser(cfg)
        got = parser.limb_argmax(head.cuda()).cpu().numpy()
        assert np.array_equal(got, want)
        clean = torch.from_numpy(synth.make_head(g, "U", seed=90 + B, B=B)).to(dtype)
        ref = c_oracle.parse_batch(clean.float().numpy(), g, n_threads=4)
        for _ in range(3):
            packed = parser.parse(clean.cuda(), input_complete=True)
        assert_packed_equals_oracle(packed.numpy(), ref, B)
    finally:
        _lib.tune(argmax_cluster=-1, argmax_cluster_ring=0)


# ------------------------------------------------------------------------------------------
# edge cases of the parse (KAT-3)
# ------------------------------------------------------------------------------------------
def _tiny():
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    graphs = (((0, 1), (1, 2)), ((0, 2), (1, 3)), ((3,), (4,)))
    return PPNConfig(K=5, E=4, insize=(128, 128), outsize=(8, 8), local_grid_size=(5, 5), directed_graphs=graphs)


def test_edge_cases_match_oracle():
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = _tiny()
    g = O.Geometry.of(cfg)
    thr = np.float32(0.15)
    imgs = []
    base = synth.make_head(g, "U", seed=1, B=1)[0]
    # 0: nothing above threshold
    a = base.copy(); a[0:g.K] = 0.1; a[g.K:2 * g.K] = 1.0; imgs.append(a)
    # 1: every delta exactly == thr: no roots (strict >)
    a = base.copy(); a[0:g.K] = thr; a[g.K:2 * g.K] = 1.0; imgs.append(a)
    # 2: one root; limb targets exactly at threshold are ACCEPTED (delta < thr breaks)
    a = base.copy(); a[0:g.K] = thr; a[g.K:2 * g.K] = 1.0; a[0, 3, 3] = 0.9; imgs.append(a)
    # 3: every limb points to the top-left window corner -> walks leave the grid near the border
    a = base.copy(); a[6 * g.K:] = 0.0; a[6 * g.K:].reshape(g.E, g.S, g.H, g.W)[:, 0] = 1.0; imgs.append(a)
    # 4: all window entries equal -> arg-max 0 everywhere
    a = base.copy(); a[6 * g.K:] = 0.5; imgs.append(a)
    # 5: identical root boxes everywhere (IoU = 1), distinct scores
    a = base.copy(); a[2 * g.K] = 0.0; a[3 * g.K] = 0.0
    a[4 * g.K] = 0.5; a[5 * g.K] = 0.5
    X, Y = np.meshgrid(np.arange(g.W, dtype=np.float32), np.arange(g.H, dtype=np.float32))
    a[2 * g.K] = (4 - X) ; a[3 * g.K] = (4 - Y)          # (x + X) constant -> same centre for every cell
    imgs.append(a)
    # 6: zero-area root boxes (w = h = 0): IoU is 0/0 = NaN -> nothing suppressed
    a = base.copy(); a[4 * g.K] = 0.0; a[5 * g.K] = 0.0; a[2 * g.K] = 0.0; a[3 * g.K] = 0.0; imgs.append(a)
    head = np.stack(imgs).astype(np.float32)
    for min_kp in (1, -1, 3):
        c2 = cfg.with_(min_num_keypoints=min_kp)
        g2 = O.Geometry.of(c2)
        ref = c_oracle.parse_batch(head, g2)
        with np.errstate(all="ignore"):
            for b in range(head.shape[0]):     # numpy twin agrees with the C twin on these too
                p = O.parse_image(head[b], g2)
                assert np.array_equal(p.part_cell, ref["part_cell"][b, :len(p.root_cell)])
        packed = PoseParser(c2).parse(torch.from_numpy(head).cuda()).numpy()
        assert_packed_equals_oracle(packed, ref, head.shape[0])
    assert ref["counts"][0, 0] == 0 and ref["counts"][1, 0] == 0 and ref["counts"][2, 0] == 1


@pytest.mark.parametrize("threads,stage", [(64, 0), (160, 1), (512, 2), (1024, 0), (256, 1), (256, 2)])
@pytest.mark.parametrize("preset,dist", [("cfg2", "D"), ("cfg4", "U")])
def test_tree_parse_cta_sizes_and_staging(preset, dist, threads, stage):
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS[preset]()
    g = O.Geometry.of(cfg)
    B = 6
    head = synth.make_head(g, dist, seed=31, B=B)
    ref = c_oracle.parse_batch(head, g, n_threads=8)
    _lib.tune(parse_threads=threads, parse_stage_all=stage)
    try:
        packed = PoseParser(cfg).parse(torch.from_numpy(head).cuda()).numpy()
    finally:
        _lib.tune(parse_threads=0, parse_stage_all=-1)
    assert_packed_equals_oracle(packed, ref, B)


def test_track_orders_that_are_not_a_tree():
    """A part reached through different limbs in different track orders: the reference's rule is
    'the later chain overwrites' (datatest.py:124); the kernel must then walk chains in order."""
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    graphs = (((0, 1), (1, 2)), ((2, 3), (3, 2)), ((0,), (4,)))      # part 2 via limb 1 and via limb 3; part 4 via limb 0 (!= part 1's)
    cfg = PPNConfig(K=5, E=4, insize=(128, 128), outsize=(8, 8), local_grid_size=(5, 5), directed_graphs=graphs,
                    detection_thresh=0.05)
    g = O.Geometry.of(cfg)
    head = synth.make_head(g, "U", seed=11, B=9)
    ref = c_oracle.parse_batch(head, g)
    assert ref["counts"][:, 2].sum() > 20
    packed = PoseParser(cfg).parse(torch.from_numpy(head).cuda()).numpy()
    assert_packed_equals_oracle(packed, ref, 9)


def test_tie_rule_is_larger_cell_first():
    """Equal root scores: the reference's order is undefined (unstable argsort); ours is pinned."""
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = _tiny().with_(min_num_keypoints=-1)
    g = O.Geometry.of(cfg)
    a = synth.make_head(g, "U", seed=2, B=1)
    a[0, 0] = 0.0; a[0, g.K] = 1.0
    a[0, 4 * g.K] = 0.01; a[0, 5 * g.K] = 0.01                  # tiny boxes: nothing suppressed
    for c in (5, 17, 40, 41):
        a[0, 0].reshape(-1)[c] = 0.5
    packed = PoseParser(cfg).parse(torch.from_numpy(a).cuda()).numpy()
    assert list(packed["root_cell"][0, :4]) == [41, 40, 17, 5]
    ref = c_oracle.parse_batch(a, g)
    assert list(ref["root_cell"][0, :4]) == [41, 40, 17, 5]


def test_threshold_rounding_rule():
    """the limb-step test `delta < detection_thresh` (datatest.py:121) is an fp32 comparison with
    float32(thr) — NumPy >= 2 semantics, which the fixtures were generated under.  At thr = 0.7 (float32(0.7) < 0.7) a
    delta exactly equal to float32(0.7) is therefore ACCEPTED (NumPy 1.x would compare in float64 and break the chain);
    the root test `delta > thr` rejects the same value in both."""
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS["cfg2"]().with_(detection_thresh=0.7)
    g = O.Geometry.of(cfg)
    thr32 = np.float32(0.7)
    assert float(thr32) < 0.7
    head = synth.make_head(g, "U", seed=4, B=1)
    K = g.K
    head[0, :K] *= np.float32(0.5)                                 # nothing above the threshold by accident
    root, tgt = 5 * g.W + 5, 5 * g.W + 6
    head[0, 0].reshape(-1)[root] = 1.0
    head[0, K].reshape(-1)[root] = 0.9                              # a root: delta 0.9 > thr
    limb0, part0 = g.graphs[0][0][0], g.graphs[0][1][0]            # first step of the first track order
    e = head[0, 6 * K:].reshape(g.E, g.S, g.H * g.W)
    e[limb0, :, root] = 0.0
    e[limb0, (g.sH // 2) * g.sW + g.sW // 2 + 1, root] = 1.0        # points one cell to the right
    head[0, part0].reshape(-1)[tgt] = 1.0
    head[0, K + part0].reshape(-1)[tgt] = thr32                      # delta == float32(thr) exactly
    ref = c_oracle.parse_batch(head, g)
    got = PoseParser(cfg).parse(torch.from_numpy(head).cuda()).numpy()
    assert_packed_equals_oracle(got, ref, 1)
    assert int(got["count"][0]) == 1 and got["part_cell"][0, 0, part0] == tgt          # accepted: not < float32(thr)
    assert bits(got["part_score"][0, 0, part0]) == bits(thr32)
    # the same value as a ROOT score is rejected by `>`
    head[0, K].reshape(-1)[root] = thr32
    got = PoseParser(cfg).parse(torch.from_numpy(head).cuda()).numpy()
    assert int(got["count"][0]) == 0


def test_capacity_limit_keeps_top_scores():
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    cfg = PPNConfig.mpii16()
    g = O.Geometry.of(cfg)
    head = synth.make_head(g, "D", seed=3, B=2)
    ref = c_oracle.parse_batch(head, g)
    packed = PoseParser(cfg, max_humans=10).parse(torch.from_numpy(head).cuda())
    a = packed.numpy()
    assert np.array_equal(a["count"], ref["counts"][:, 2]) and a["count"].min() > 10
    assert np.array_equal(a["part_cell"][:, :10], ref["part_cell"][:, :10])
    with pytest.raises(RuntimeError):
        packed.to_lists()


def test_pack_dense_entries_roundtrip():
    """ppn_pack_humans: fixed-stride result -> dense (human, part) entries (what the multi-GPU gather ships)."""
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser, entries_to_packed, unpack_entries
    for cfg, dist in ((PPNConfig.mpii16(), "U"), (PPNConfig.coco18(), "D")):
        g = O.Geometry.of(cfg)
        B = 21
        head = synth.make_head(g, dist, seed=5, B=B)
        ref = c_oracle.parse_batch(head, g, n_threads=8)
        parser = PoseParser(cfg)
        out = parser.parse(torch.from_numpy(head).cuda())
        want_entries = np.array([(ref["part_cell"][b, :ref["counts"][b, 2]] >= 0).sum() for b in range(B)])
        total = int(want_entries.sum())
        for cap in (total + 10, total, total // 2):
            buf = parser.pack(out, cap)
            _, offs = parser.packed_layout(B, cap)
            rec = unpack_entries(buf.cpu(), B, cap, offs)
            assert rec["total"] == total and rec["overflow"] == (total > cap)
            assert np.array_equal(rec["count"], ref["counts"][:, 2])
            assert np.array_equal(rec["entries"], want_entries)
            for b in range(B):
                if rec["start"][b] + rec["entries"][b] > cap:
                    continue                                       # (partly) dropped: overflow was flagged
                n = int(ref["counts"][b, 2])
                pc, ps, pb = entries_to_packed(rec, b, cfg.K)
                assert np.array_equal(pc, ref["part_cell"][b, :n])
                assert np.array_equal(bits(ps), bits(ref["part_score"][b, :n]))
                assert np.array_equal(bits(pb), bits(ref["part_box"][b, :n]))


@pytest.mark.parametrize("skip_slots", [False, True])
def test_parse_dense_written_by_the_parse_kernel(fused, skip_slots):
    """ppn_parse_dense: the dense (human, part) entries straight from the parse kernel (two-kernel path;
    image blocks in any order, header.start says where) or via the pack kernels (three-kernel path) —
    the same records either way, overlapped calls into rotating buffers included."""
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser, entries_to_packed, unpack_entries
    for cfg, dist, B in ((PPNConfig.mpii16(), "U", 37), (PPNConfig.coco18(), "D", 9), (PPNConfig.mpii16(), "S", 300)):
        g = O.Geometry.of(cfg)
        heads = [synth.make_head(g, dist, seed=60 + i, B=B) for i in range(2)]
        refs = [c_oracle.parse_batch(h, g, n_threads=8) for h in heads]
        devs = [torch.from_numpy(h).cuda() for h in heads]
        parser = PoseParser(cfg)
        totals = [int(sum((r["part_cell"][b, :r["counts"][b, 2]] >= 0).sum() for b in range(B))) for r in refs]
        for cap in (max(totals) + 7, max(totals) // 2):
            nbytes, offs = parser.packed_layout(B, cap)
            n_calls = 9
            bufs = [torch.full((nbytes,), 0xAB, dtype=torch.uint8, device="cuda") for _ in range(n_calls)]
            outs = [parser.alloc_output(B) for _ in range(n_calls)]
            for o in outs:
                o.part_cell.fill_(-7)
            torch.cuda.synchronize()
            for i in range(n_calls):
                parser.parse(devs[i % 2], out=outs[i], input_complete=(i != 4), dense=bufs[i], cap_entries=cap, skip_slots=skip_slots)
            torch.cuda.synchronize()
            for i in range(n_calls):
                ref, total = refs[i % 2], totals[i % 2]
                rec = unpack_entries(bufs[i].cpu(), B, cap, offs)
                want_entries = np.array([(ref["part_cell"][b, :ref["counts"][b, 2]] >= 0).sum() for b in range(B)])
                assert rec["total"] == total and rec["overflow"] == (total > cap)
                assert np.array_equal(rec["count"], ref["counts"][:, 2])
                assert np.array_equal(rec["entries"], want_entries)
                assert np.array_equal(outs[i].count.cpu().numpy(), ref["counts"][:, 2])
                spans = []
                for b in range(B):
                    if rec["start"][b] + rec["entries"][b] > cap:
                        continue                                   # dropped: overflow was flagged
                    spans.append((int(rec["start"][b]), int(rec["start"][b] + rec["entries"][b])))
                    n = int(ref["counts"][b, 2])
                    pc, ps, pb = entries_to_packed(rec, b, cfg.K)
                    assert np.array_equal(pc, ref["part_cell"][b, :n])
                    assert np.array_equal(bits(ps), bits(ref["part_score"][b, :n]))
                    assert np.array_equal(bits(pb), bits(ref["part_box"][b, :n]))
                spans.sort()
                assert all(a[1] <= b2[0] for a, b2 in zip(spans, spans[1:])), "image blocks overlap"
                if total <= cap:
                    assert sum(e - s0 for s0, e in spans) == total
                if not skip_slots:
                    assert_packed_equals_oracle(outs[i].numpy(), ref, B)


def test_two_kernel_chain_cut_into_sub_batches():
    """A batch whose parse CTAs do not all fit beside the arg-max ring (16x16 grid, 700 images) forced onto
    the two-kernel chain: it is cut into two sub-batches (4 launches), with and without the dense entry
    buffer, overlapped and not — results as the oracle's."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser, entries_to_packed, unpack_entries
    cfg = PRESETS["cfg3"]()
    g = O.Geometry.of(cfg)
    B = 700
    head = synth.make_head(g, "D", seed=321, B=B)
    ref = c_oracle.parse_batch(head, g, n_threads=8)
    dev = torch.from_numpy(head).cuda()
    _lib.tune(parse_fused=1)
    try:
        parser = PoseParser(cfg)
        assert parser.launches_per_parse(B) == 4 and parser.launches_per_parse(500) == 2
        total = int(sum((ref["part_cell"][b, :ref["counts"][b, 2]] >= 0).sum() for b in range(B)))
        nbytes, offs = parser.packed_layout(B, total + 5)
        outs = [parser.alloc_output(B) for _ in range(4)]
        bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(4)]
        torch.cuda.synchronize()
        for i in range(4):
            parser.parse(dev, out=outs[i], input_complete=(i != 2), dense=bufs[i] if i % 2 else None, cap_entries=total + 5)
        torch.cuda.synchronize()
        for i in range(4):
            assert_packed_equals_oracle(outs[i].numpy(), ref, B)
        for i in (1, 3):
            rec = unpack_entries(bufs[i].cpu(), B, total + 5, offs)
            assert rec["total"] == total and not rec["overflow"]
            assert np.array_equal(rec["count"], ref["counts"][:, 2])
            for b in range(0, B, 7):
                n = int(ref["counts"][b, 2])
                pc, ps, pb = entries_to_packed(rec, b, cfg.K)
                assert np.array_equal(pc, ref["part_cell"][b, :n])
                assert np.array_equal(bits(pb), bits(ref["part_box"][b, :n]))
    finally:
        _lib.tune(parse_fused=-1)


def _gatherer_worker(rank, port, tmp):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1")
    import torch.distributed as dist
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser, entries_to_packed
    from pytorch_pose_proposal_network_b200.sharded import PoseGatherer
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        cfg = PPNConfig.mpii16()
        g = O.Geometry.of(cfg)
        B, steps, gs = 48, 11, 4                                   # 2 full groups + a partial one
        heads = [synth.make_head(g, "U", seed=40 + i, B=B) for i in range(3)]
        refs = [c_oracle.parse_batch(h, g, n_threads=4) for h in heads]
        devs = [torch.from_numpy(h).cuda() for h in heads]
        parser = PoseParser(cfg)
        gat = PoseGatherer(parser, B, cap_entries=B * 120, group_steps=gs)
        outs = [parser.alloc_output(B) for _ in range(2)]
        for i in range(steps):
            gat.parse(devs[i % 3], out=outs[i % 2], input_complete=True)
        gat.finish()
        torch.cuda.synchronize()
        ok = True
        for i in range(((steps - 1) // gs - 1) * gs, steps):       # the two most recent groups are still held
            rec, ref = gat.records_of(0, step_back=steps - 1 - i), refs[i % 3]
            ok &= not rec["overflow"] and np.array_equal(rec["count"], ref["counts"][:, 2])
            for b in range(0, B, 5):
                n = int(ref["counts"][b, 2])
                pc, ps, pb = entries_to_packed(rec, b, cfg.K)
                ok &= np.array_equal(pc, ref["part_cell"][b, :n]) and np.array_equal(bits(pb), bits(ref["part_box"][b, :n]))
        open(os.path.join(tmp, "ok" if ok else "bad"), "w").close()
    finally:
        dist.destroy_process_group()


def test_pose_gatherer_single_rank_nccl(tmp_path):
    """sharded.PoseGatherer.parse end to end on one GPU (NCCL group of one): dense records written by the
    parse kernel into rotating group buffers, async all_gather per group, partial last group flushed."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + 17
    mp.spawn(_gatherer_worker, args=(port, str(tmp_path)), nprocs=1, join=True)
    assert os.listdir(tmp_path) == ["ok"]


def _peer_gatherer_worker(rank, port, tmp, control, mode):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1")
    import torch.distributed as dist
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser, entries_to_packed
    from pytorch_pose_proposal_network_b200.sharded import PeerPoseGatherer
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        cfg = PPNConfig.mpii16()
        g = O.Geometry.of(cfg)
        B, steps, slots = 48, 21, 8                                # the ring of slots wraps twice
        heads = [synth.make_head(g, "U", seed=60 + i, B=B) for i in range(3)]
        refs = [c_oracle.parse_batch(h, g, n_threads=4) for h in heads]
        devs = [torch.from_numpy(h).cuda() for h in heads]
        parser = PoseParser(cfg)
        gat = PeerPoseGatherer(parser, B, B * 120, slots=slots, notify_every=3, mode=mode, control=control)
        outs = [parser.alloc_output(B) for _ in range(2)]
        for i in range(steps):
            gat.parse(devs[i % 3], out=outs[i % 2], input_complete=True)
        gat.finish()
        torch.cuda.synchronize()
        gat.check_landed()
        ok = True
        held = slots if control == "flags" else slots - 2 * 3 + 1
        for i in range(steps - held, steps):                       # the slots still hold the most recent steps
            rec, ref = gat.records_of(0, step_back=steps - 1 - i), refs[i % 3]
            ok &= not rec["overflow"] and np.array_equal(rec["count"], ref["counts"][:, 2])
            for b in range(0, B, 5):
                n = int(ref["counts"][b, 2])
                pc, ps, pb = entries_to_packed(rec, b, cfg.K)
                ok &= np.array_equal(pc, ref["part_cell"][b, :n]) and np.array_equal(bits(pb), bits(ref["part_box"][b, :n]))
        for i in range(steps, steps + 4):                         # a second run on the same gatherer: the counters carry the run number
            gat.parse(devs[i % 3], out=outs[i % 2], input_complete=True)
        gat.finish()
        torch.cuda.synchronize()
        gat.check_landed()
        rec, ref = gat.records_of(0), refs[(steps + 3) % 3]
        ok &= not rec["overflow"] and np.array_equal(rec["count"], ref["counts"][:, 2])
        gat.close()
        open(os.path.join(tmp, "ok" if ok else "bad"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("control,mode", [("flags", "store"), ("flags", "copy"), ("nccl", "store")])
def test_peer_pose_gatherer_single_rank(tmp_path, control, mode):
    """sharded.PeerPoseGatherer on one GPU (process group of one, the 'root' buffer is local): records stored by the
    parse kernel (or copied) into the slot ring, landing announced by counters in the buffer (`flags`) or by NCCL."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + 31
    mp.spawn(_peer_gatherer_worker, args=(port, str(tmp_path), control, mode), nprocs=1, join=True)
    assert os.listdir(tmp_path) == ["ok"]


def test_peer_wait_times_out_instead_of_hanging():
    """ppn_peer_wait on a counter nobody posts: gives up after the timeout and raises the flag; after ppn_peer_post
    the same wait passes."""
    from pytorch_pose_proposal_network_b200 import _lib
    lib = _lib.lib()
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.ppn_peer_wait(counters.data_ptr(), 4, 3, 20, flag.data_ptr(), st), "ppn_peer_wait")
    torch.cuda.synchronize()
    assert int(flag.item()) == 1
    flag.zero_()
    for r in range(4):
        _lib.check(lib.ppn_peer_post(counters.data_ptr() + 8 * r, 3 + r, st), "ppn_peer_post")
    _lib.check(lib.ppn_peer_wait(counters.data_ptr(), 4, 3, 2000, flag.data_ptr(), st), "ppn_peer_wait")
    torch.cuda.synchronize()
    assert int(flag.item()) == 0 and counters.tolist() == [3, 4, 5, 6]
    assert lib.ppn_peer_post(None, 1, st) != 0 and lib.ppn_peer_wait(None, 4, 0, 1, None, st) != 0       # bad arguments are refused


def test_bad_arguments_raise():
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PPNConfig.mpii16()
    p = PoseParser(cfg)
    with pytest.raises(ValueError):
        p.parse(torch.zeros(1, cfg.C + 1, cfg.H, cfg.W, device="cuda"))
    with pytest.raises(ValueError):
        p.parse(torch.zeros(1, cfg.C, cfg.H, cfg.W, device="cuda", dtype=torch.float64))
    shape = p.c.shape(1)
    hs = _lib.PPNHumans(0, 0, 0, 0, 0, 4)
    rc = _lib.lib().ppn_parse(None, C.byref(shape), C.byref(p.c.params), C.byref(hs), None, 0, None)
    assert rc == -1 and b"bad argument" in _lib.lib().ppn_strerror(rc)
    out = p.alloc_output(1)
    hs = p._humans_struct(out)
    ws = torch.empty(256, dtype=torch.uint8, device="cuda")
    head = torch.zeros(1, cfg.C, cfg.H, cfg.W, device="cuda")
    rc = _lib.lib().ppn_parse(head.data_ptr(), C.byref(shape), C.byref(p.c.params), C.byref(hs), ws.data_ptr(), 256, None)
    assert rc == -3
    with pytest.raises(ValueError):
        PoseParser(PPNConfig.mpii16(outsize=(40, 40), insize=(640, 640)))


# ------------------------------------------------------------------------------------------
# BASELINE.json's full sizes: size-independent properties + sampled oracle check
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("preset,B,dist", [("cfg2", 512, "U"), ("cfg3", 1024, "D"), ("cfg4", 256, "U")])
def test_full_size_properties(preset, B, dist):
    from pytorch_pose_proposal_network_b200.config import PRESETS
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PRESETS[preset]()
    g = O.Geometry.of(cfg)
    from tests.gpu_inputs import device_head, distinct_root_scores
    gen = torch.Generator(device="cuda").manual_seed(1234)
    head, host_all = device_head(g, dist, seed=1234, B=B)      # equal root scores, if any, already moved apart
    parser = PoseParser(cfg)
    first = {k: v.copy() for k, v in parser.parse(head).numpy().items()}
    # determinism / idempotence: same input, same bytes (compare only the valid slots)
    second = parser.parse(head).numpy()
    assert np.array_equal(first["count"], second["count"])
    # permutation equivariance: images are independent, so parse(head[perm]) == parse(head)[perm]
    perm = torch.randperm(B, device="cuda", generator=gen)
    shuffled = parser.parse(head[perm].contiguous()).numpy()
    pc = perm.cpu().numpy()
    assert np.array_equal(shuffled["count"], first["count"][pc])
    for i in range(0, B, max(1, B // 64)):
        n = int(shuffled["count"][i])
        assert np.array_equal(shuffled["part_cell"][i, :n], first["part_cell"][pc[i], :n])
        assert np.array_equal(bits(shuffled["part_box"][i, :n]), bits(first["part_box"][pc[i], :n]))
    # structural invariants of every human
    for b in range(0, B, max(1, B // 32)):
        n = int(first["count"][b])
        cells = first["part_cell"][b, :n]
        assert (cells[:, 0] == first["root_cell"][b, :n]).all()
        assert ((cells >= -1) & (cells < cfg.HW)).all()
        assert ((cells[:, 1:] >= 0).sum(1) >= cfg.min_num_keypoints).all()
        sc = first["part_score"][b, :n]
        assert (sc[cells >= 0] >= np.float32(cfg.detection_thresh)).all()
        assert (np.diff(sc[:, 0]) <= 0).all()                 # descending root score
    # a sample of images against the C restatement, bit for bit
    pick = np.linspace(0, B - 1, 12).astype(int)
    sub = np.ascontiguousarray(host_all[pick])
    assert distinct_root_scores(sub, g) == 0 and synth.root_scores_distinct(sub, g)
    ref = c_oracle.parse_batch(sub, g, n_threads=8)
    sample = {k: v[pick] for k, v in first.items()}
    assert_packed_equals_oracle(sample, ref, len(pick))


# ------------------------------------------------------------------------------------------
# fuzz: random geometries, skeletons and thresholds against the C restatement
# ------------------------------------------------------------------------------------------
def _random_tree(rng, K):
    """Random rooted tree over parts 0..K-1 as track orders (root-to-leaf chains) + limb list."""
    parent = [-1] + [int(rng.integers(0, k)) for k in range(1, K)]
    limb_of = {k: k - 1 for k in range(1, K)}                 # limb k-1 joins parent[k] -> k
    children = {k: [] for k in range(K)}
    for k in range(1, K):
        children[parent[k]].append(k)
    leaves = [k for k in range(K) if not children[k]]
    graphs = []
    for leaf in leaves:
        path = []
        k = leaf
        while k != 0:
            path.append(k)
            k = parent[k]
        path.reverse()
        if path:
            graphs.append((tuple(limb_of[t] for t in path), tuple(path)))
    return tuple(graphs), K - 1


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_random_geometry(seed):
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    rng = np.random.default_rng(1000 + seed)
    K = int(rng.integers(2, 20))
    graphs, E = _random_tree(rng, K)
    W, H = int(rng.integers(2, 27)), int(rng.integers(2, 27))
    while W * H > 1024:
        W, H = int(rng.integers(2, 27)), int(rng.integers(2, 27))
    s = int(rng.choice([1, 3, 5, 7, 9, 13]))
    cell = int(rng.choice([8, 16, 32]))
    cfg = PPNConfig(K=K, E=E, insize=(W * cell, H * cell), outsize=(W, H), local_grid_size=(s, s),
                    directed_graphs=graphs, detection_thresh=float(rng.choice([0.05, 0.15, 0.3, 0.6])),
                    nms_thresh=float(rng.choice([0.1, 0.3, 0.5, 0.9])), min_num_keypoints=int(rng.choice([-1, 1, 2])))
    g = O.Geometry.of(cfg)
    B = int(rng.integers(1, 9))
    head = synth.make_head(g, str(rng.choice(["U", "R", "D", "S"])), seed=seed, B=B)
    if not synth.root_scores_distinct(head, g, cfg.detection_thresh):
        pytest.skip("duplicate root scores in this draw")
    ref = c_oracle.parse_batch(head, g, n_threads=4)
    parser = PoseParser(cfg)
    dev = torch.from_numpy(head).cuda()
    assert_packed_equals_oracle(parser.parse(dev).numpy(), ref, B)
    assert_packed_equals_oracle(parser.parse(dev, input_complete=True).numpy(), ref, B)
    want = np.stack([c_oracle.limb_argmax(img, g) for img in head]).astype(np.uint16) if E else None
    if E:
        assert np.array_equal(parser.limb_argmax(dev).cpu().numpy(), want)
