"""Training-target encoder (the inverse of the parser, dataset.py:98-185).

CPU: the numpy restatement against outputs of the reference's own ``__getitem__`` (tests/golden/encode,
made by oracle/make_golden_encode.py), and its round trip through the parser's restatement.
GPU: ``ppn_encode_targets`` against the same fixtures and against the restatement on fresh random
annotation sets — bit for bit on all ten tensors.
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import encode_gt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encode")
NAMES = ("delta", "weight", "weight_ij", "tx", "ty", "tx_half", "ty_half", "tw", "th", "te")


def cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load(name):
    fx = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    meta = json.loads(str(fx["meta"]))
    want = {}
    for nm in NAMES:
        if nm in ("te", "weight_ij"):
            full = np.full(int(np.prod(fx[nm + "_shape"])), 0.0 if nm == "te" else 0.0005, np.float32)
            full[fx[nm + "_ones"]] = 1.0
            want[nm] = full.reshape(tuple(fx[nm + "_shape"]))
        else:
            want[nm] = fx[nm]
    return meta, fx, want


def edges_of(K):
    from pytorch_pose_proposal_network_b200 import config as pcfg
    return pcfg.EDGES if K == 18 else pcfg.EDGES_16


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_fixtures_exist():
    assert len(cases()) >= 4


@pytest.mark.parametrize("name", cases())
def test_restatement_matches_reference_outputs(name):
    meta, fx, want = load(name)
    off = fx["person_off"]
    for b in range(len(off) - 1):
        sl = slice(off[b], off[b + 1])
        got = encode_gt.encode_targets(fx["keypoints"][sl], fx["bbox"][sl], fx["visible"][sl].astype(bool), fx["size"][sl],
                                       meta["K"], edges_of(meta["K"]), meta["insize"], meta["outsize"], meta["window"])
        for nm, g in zip(NAMES, got):
            assert np.array_equal(bits(g), bits(want[nm][b])), (name, b, nm)


def test_encoded_targets_parse_back():
    """The reference's only self-check (datatest.py:403-412): targets fed to the parser give the people back."""
    from oracle import ppn_oracle as O
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    cfg = PPNConfig.reference_native(outsize=(12, 12), local_grid_size=(9, 9))
    g = O.Geometry.of(cfg)
    K = cfg.K
    kp = np.zeros((2, K - 1, 2), np.float32)
    vis = np.zeros((2, K - 1), bool)
    # two people, three connected parts each (instance -> neck (15) -> thorax (13)), well apart
    for p, (x0, y0) in enumerate([(60.0, 70.0), (300.0, 250.0)]):
        for k, (dx, dy) in ((15, (8.0, 40.0)), (13, (12.0, 75.0))):
            kp[p, k - 1] = (x0 + dx, y0 + dy)
            vis[p, k - 1] = True
    bbox = np.array([[60.0, 70.0, 50.0, 60.0], [300.0, 250.0, 40.0, 70.0]])
    t = encode_gt.encode_targets(kp, bbox, vis, [20.0, 24.0], K, edges_of(K), cfg.insize, cfg.outsize, cfg.local_grid_size)
    delta, tx, ty, tw, th, te = t[0], t[3], t[4], t[7], t[8], t[9]
    head = np.concatenate([delta, np.ones_like(delta), tx, ty, tw, th, te.reshape(-1, cfg.H, cfg.W)]).astype(np.float32)
    head[K + 0] = 1.0
    head[K, 70 // 32, 60 // 32] = 0.9                                   # distinct root scores (conf of part 0)
    res = O.parse_image(head, g)
    humans, _ = O.humans_as_dicts(res)
    assert len(humans) == 2 and all(set(h) == {0, 15, 13} for h in humans)


def test_flatten_samples_matches_the_reference_zip():
    """Host side of TargetEncoder: the reference's encoder loop zips bbox, keypoints, is_visible and size
    (dataset.py:108), i.e. stops at the shortest — aug.py:115-116 leaves ONE zero person in `keypoints` when no
    box survived, which must not become a person here either."""
    from pytorch_pose_proposal_network_b200.dataset import flatten_samples
    K = 18
    rng = np.random.default_rng(3)
    kp, bb, vis, size = encode_gt.random_people(rng, 3, K, (384, 384))
    empty = dict(keypoints=np.zeros((1, K - 1, 2), np.float32), bbox=np.zeros((0, 4)), is_visible=[], size=[])
    full = dict(keypoints=torch.from_numpy(kp), bbox=torch.from_numpy(bb), is_visible=vis, size=size)
    off, fb, fk, fv, fs = flatten_samples([empty, full, empty], K)
    assert off.tolist() == [0, 0, 3, 3] and off.dtype == np.int32
    assert fb.dtype == np.float64 and np.array_equal(fb, bb)
    assert fk.dtype == np.float32 and np.array_equal(fk, kp)
    assert fv.dtype == np.uint8 and np.array_equal(fv.astype(bool), np.asarray(vis))
    assert fs.dtype == np.float64 and np.array_equal(fs, np.asarray(size))
    off, fb, fk, fv, fs = flatten_samples([], K)
    assert off.tolist() == [0] and fb.shape == (0, 4) and fk.shape == (0, K - 1, 2) and fv.shape == (0, K - 1)


def test_encoder_has_no_cpu_path():
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.dataset import TargetEncoder
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    with pytest.raises(RuntimeError):
        TargetEncoder(PPNConfig.reference_native())


# ------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def _encoder(meta):
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.dataset import TargetEncoder
    K = meta["K"]
    base = PPNConfig.reference_native if K == 18 else PPNConfig.mpii16
    cfg = base(insize=tuple(meta["insize"]), outsize=tuple(meta["outsize"]), local_grid_size=tuple(meta["window"]))
    return TargetEncoder(cfg)


@pytest.fixture(params=[1, 0], ids=["sweep", "per_image"])
def sweep(request):
    """Both ways of writing the two limb tensors: the address-ordered persistent sweep and one CTA per image part."""
    from pytorch_pose_proposal_network_b200 import _lib
    _lib.tune(encode_sweep=request.param)
    yield request.param
    _lib.tune(encode_sweep=1)


@gpu
@pytest.mark.parametrize("name", cases())
def test_gpu_encoder_matches_reference_outputs(name, sweep):
    meta, fx, want = load(name)
    enc = _encoder(meta)
    dev = [torch.from_numpy(np.ascontiguousarray(fx[k])).cuda() for k in ("person_off", "bbox", "keypoints", "visible", "size")]
    out = enc.alloc(len(fx["person_off"]) - 1)
    for t in out.as_list():
        t.fill_(float("nan"))                                          # every element must be written
    enc.encode_flat(*dev, out=out)
    torch.cuda.synchronize()
    for nm in NAMES:
        assert np.array_equal(bits(getattr(out, nm).cpu().numpy()), bits(want[nm])), (name, nm)


@gpu
@pytest.mark.parametrize("seed", range(6))
def test_gpu_encoder_matches_restatement_random(seed, sweep):
    """Fresh annotation sets in the reference's sample format through TargetEncoder.encode: crowded cells
    (later people overwrite earlier ones), points outside the image on every side, empty images, batch sizes
    around the CTA-splitting thresholds, grids whose width is not a multiple of four (scalar store path)."""
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.dataset import TargetEncoder
    rng = np.random.default_rng(1000 + seed)
    geo = [((384, 384), (12, 12), (9, 9)), ((384, 384), (24, 24), (21, 21)), ((416, 416), (13, 13), (7, 7)),
           ((768, 768), (24, 24), (11, 11)), ((384, 256), (12, 8), (5, 5)), ((320, 320), (10, 10), (3, 3))][seed]
    K = 18 if seed % 2 == 0 else 16
    base = PPNConfig.reference_native if K == 18 else PPNConfig.mpii16
    cfg = base(insize=geo[0], outsize=geo[1], local_grid_size=geo[2])
    B = [1, 3, 40, 7, 700, 149][seed]
    samples, raw = [], []
    for b in range(B):
        n = int(rng.integers(0, 30 if b % 5 == 0 else 6))
        kp, bb, vis, size = encode_gt.random_people(rng, n, K, geo[0], spread=1.6 if b % 3 == 0 else 1.1)
        raw.append((kp, bb, vis, size))
        samples.append(dict(keypoints=torch.from_numpy(kp), bbox=torch.from_numpy(np.asarray(bb).reshape(-1, 4)), is_visible=vis, size=size))
    out = TargetEncoder(cfg).encode(samples)
    torch.cuda.synchronize()
    got = {nm: getattr(out, nm).cpu().numpy() for nm in NAMES}
    for b in range(0, B, max(1, B // 25)):
        kp, bb, vis, size = raw[b]
        want = encode_gt.encode_targets(kp, bb, vis, size, K, edges_of(K), *geo)
        for nm, w in zip(NAMES, want):
            assert np.array_equal(bits(got[nm][b]), bits(w)), (seed, b, nm)


@gpu
def test_gpu_encoder_rejects_what_the_reference_cannot_do():
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.dataset import TargetEncoder
    enc = TargetEncoder(PPNConfig.reference_native(outsize=(12, 12), local_grid_size=(9, 7)))
    with pytest.raises(_lib.PPNError):
        enc.encode([dict(keypoints=np.zeros((0, 17, 2), np.float32), bbox=np.zeros((0, 4)), is_visible=[], size=[])])
    with pytest.raises(ValueError):
        TargetEncoder(PPNConfig.reference_native(), edges=[[0, 1]])


@gpu
def test_gpu_encode_then_gpu_parse_round_trip_full_batch():
    """Both directions together at BASELINE batch size: 512 annotation sets -> ppn_encode_targets -> a head tensor
    assembled on the device exactly as datatest.py:405-412 feeds the parser (resp = delta targets, conf = 1, limb block
    = te) -> ppn_parse, against the numpy encoder restatement followed by the C parser restatement.  The inputs are
    the parser's worst case for ties: every score is exactly 1, every empty limb window is all zeros (arg-max 0)."""
    from oracle import c_oracle, ppn_oracle as O
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.dataset import TargetEncoder
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    cfg = PPNConfig.mpii16()                                   # K16 / E15, 384x384, 12x12 grid, 9x9 window
    g = O.Geometry.of(cfg)
    K, B = cfg.K, 512
    rng = np.random.default_rng(2024)
    raw = [encode_gt.random_people(rng, int(rng.integers(0, 13)), K, cfg.insize, spread=1.05) for _ in range(B)]
    samples = [dict(keypoints=kp, bbox=bb, is_visible=vis, size=size) for kp, bb, vis, size in raw]
    t = TargetEncoder(cfg).encode(samples)
    head = torch.cat([t.delta, torch.ones_like(t.delta), t.tx, t.ty, t.tw, t.th,
                      t.te.reshape(B, cfg.E * cfg.S, cfg.H, cfg.W)], dim=1).contiguous()
    packed = PoseParser(cfg).parse(head).numpy()
    want_head = np.empty((B, cfg.C, cfg.H, cfg.W), np.float32)
    for b, (kp, bb, vis, size) in enumerate(raw):
        w = encode_gt.encode_targets(kp, bb, vis, size, K, edges_of(K), cfg.insize, cfg.outsize, cfg.local_grid_size)
        want_head[b] = np.concatenate([w[0], np.ones_like(w[0]), w[3], w[4], w[7], w[8], w[9].reshape(-1, cfg.H, cfg.W)])
    assert np.array_equal(bits(head.cpu().numpy()), bits(want_head))
    ref = c_oracle.parse_batch(want_head, g, n_threads=8)
    assert np.array_equal(packed["count"], ref["counts"][:, 2]) and int(packed["count"].sum()) > B // 2
    for b in range(B):
        n = int(ref["counts"][b, 2])
        assert np.array_equal(packed["part_cell"][b, :n], ref["part_cell"][b, :n]), b
        assert np.array_equal(bits(packed["part_box"][b, :n]), bits(ref["part_box"][b, :n])), b
        assert np.array_equal(bits(packed["part_score"][b, :n]), bits(ref["part_score"][b, :n])), b
