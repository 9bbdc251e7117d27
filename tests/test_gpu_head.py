"""Parity of the fused network head (csrc/ppn_head.cu: tcgen05 1x1-convolution GEMM + sigmoid + limb arg-max,
then the fused parse) — SURVEY §8f row 1, /root/reference/model.py:85,133-136.

Parity definition:
  * logits: the kernel can emit its convolution output; it must match torch.nn.functional.conv2d evaluated in
    fp32 (TF32 off) within |d| <= 4e-3 + 4e-3 * |ref| — the kernel multiplies in TF32 (10-bit mantissas, what
    cuDNN does for the reference under PyTorch's defaults) and accumulates 512 products in fp32;
  * everything after the logits is EXACT: the emitted head tensor is sigmoid(logits) and the decode planes, the
    arg-max map, the NMS survivors, the human assignment, boxes and scores must equal, bit for bit, what the
    oracle computes from that emitted tensor.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle, ppn_oracle as O
from tests.test_gpu_parity import assert_packed_equals_oracle, bits, cfg_of, parser_for

pytestmark = pytest.mark.gpu

ATOL, RTOL = 4e-3, 4e-3
# 16-bit operands: against the fp32 convolution of the UNROUNDED operands — fp16 keeps TF32's 10
# mantissa bits, bf16 keeps 7; against the convolution of the operands rounded to 16 bits only the fp32 accumulation
# of Cin products is left
TOL16 = {"f16": (4e-3, 4e-3), "bf16": (3e-2, 3e-2)}
TOL_ACC = (1e-4, 1e-4)
DT16 = {"f16": torch.float16, "bf16": torch.bfloat16}


def geometry(name):
    from pytorch_pose_proposal_network_b200.config import PRESETS
    return O.Geometry.of(PRESETS[name]())


def make_layer(g, B, Cin, seed, device="cuda"):
    """Synthetic input of the last layer and conv3 parameters, generated on the device: activations
    like a leaky-ReLU output, kaiming-normal weights as model.py:97-99 initialises them, a bias that spreads the
    part responses so that roots and limbs exist."""
    gen = torch.Generator(device=device).manual_seed(seed)
    C = 6 * g.K + g.S * g.E
    feat = torch.randn(B, Cin, g.H, g.W, device=device, generator=gen)
    feat = torch.where(feat > 0, feat, 0.1 * feat).contiguous()
    weight = (torch.randn(C, Cin, device=device, generator=gen) * (2.0 / (1.01 * Cin)) ** 0.5).contiguous()
    bias = torch.randn(C, device=device, generator=gen) * 0.5
    bias[:2 * g.K] += 1.0                                  # resp, conf: enough cells above the detection threshold
    return feat, weight, bias.contiguous()


def conv_fp32(feat, weight, bias):
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return torch.nn.functional.conv2d(feat.double(), weight.double()[:, :, None, None], bias.double()).float()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def check_against_emitted(g, parser, dec, amax, logits, head):
    """exact checks of everything downstream of the logits"""
    K6 = 6 * g.K
    B = head.shape[0]
    h = head.cpu().numpy()
    assert np.array_equal(bits(dec.cpu().numpy()), bits(h[:, :K6])), "decode planes differ from the emitted head"
    limbs = h[:, K6:].reshape(B, g.E, g.S, g.H * g.W)
    want = np.argmax(limbs, axis=2).astype(np.uint16).reshape(B, g.E, g.H, g.W)     # numpy: first maximum, first NaN
    got = amax.cpu().numpy()
    assert np.array_equal(got, want), f"{int((got != want).sum())} arg-max entries differ"
    # the emitted head is the sigmoid of the emitted logits (torch's CUDA expression; 1 ulp allowed for libm differences)
    ref = torch.sigmoid(logits)
    d = (ref.view(torch.int32) - head.view(torch.int32)).abs()
    finite = torch.isfinite(ref) & torch.isfinite(head)
    assert int(d[finite].max()) <= 1, f"sigmoid differs from torch.sigmoid by {int(d[finite].max())} ulp"


@pytest.mark.parametrize("name,B,Cin", [("cfg1", 3, 64), ("cfg1", 5, 512), ("cfg3", 2, 512), ("cfg4", 1, 512), ("native", 2, 512)])
def test_head_kernel_logits_and_epilogue(name, B, Cin):
    g = geometry(name)
    parser = parser_for(g)
    feat, weight, bias = make_layer(g, B, Cin, seed=50 + B)
    dec, amax, logits, head = parser.head_gemm_argmax(feat, weight, bias, emit=True)
    torch.cuda.synchronize()
    ref = conv_fp32(feat, weight, bias)
    err = (logits - ref).abs()
    tol = ATOL + RTOL * ref.abs()
    assert bool((err <= tol).all()), f"max |logit error| {float(err.max()):.3e} (limit {ATOL} + {RTOL}|ref|)"
    check_against_emitted(g, parser, dec, amax, logits, head)
    # without the emit buffers the kernel must produce the same planes and map
    dec2, amax2, _, _ = parser.head_gemm_argmax(feat, weight, bias, emit=False)
    torch.cuda.synchronize()
    assert torch.equal(dec2, dec) and torch.equal(amax2.view(torch.int16), amax.view(torch.int16))


@pytest.mark.parametrize("name,B", [("cfg1", 7), ("cfg2", 64), ("cfg3", 8), ("native", 3)])
def test_head_parse_matches_oracle_on_emitted_head(name, B):
    g = geometry(name)
    parser = parser_for(g)
    feat, weight, bias = make_layer(g, B, 512, seed=77)
    packed, logits, head = parser.parse_features(feat, weight, bias, emit=True)
    torch.cuda.synchronize()
    ref = c_oracle.parse_batch(head.cpu().numpy(), g)
    assert int(ref["counts"][:, 2].sum()) > 0, "degenerate test input: no humans"
    assert_packed_equals_oracle(packed.numpy(), ref, B)
    # ... and equals the two-step path (materialised head tensor -> ppn_parse), and the run without emit buffers
    two_step = parser.parse(head, out=parser.alloc_output(B)).numpy()
    plain = parser.parse_features(feat, weight, bias, out=parser.alloc_output(B)).numpy()
    torch.cuda.synchronize()
    for other in (two_step, plain):
        assert np.array_equal(other["count"], packed.numpy()["count"])
        assert_packed_equals_oracle(other, ref, B)


def test_head_exact_logits_ties_nan_inf():
    """Logits under full control: zero weights, so logit == bias exactly.  Neighbouring logits that share a sigmoid
    value must resolve to the FIRST one (numpy arg-max on the sigmoid values), NaN wins, +-inf behave."""
    g = geometry("cfg1")
    parser = parser_for(g)
    B, Cin = 2, 32
    C, K6 = 6 * g.K + g.S * g.E, 6 * g.K
    feat = torch.randn(B, Cin, g.H, g.W, device="cuda")
    weight = torch.zeros(C, Cin, device="cuda")
    rng = np.random.default_rng(5)
    bias = rng.normal(0, 1, C).astype(np.float32)
    lim = bias[K6:].reshape(g.E, g.S)
    up = lambda v, n: (np.float32(v).view(np.uint32) + np.uint32(n)).view(np.float32)      # n ulps above a positive float
    lim[0, :] = -3.0; lim[0, 10] = 1.5; lim[0, 40] = up(1.5, 1)          # 1 ulp apart: same sigmoid -> index 10
    lim[1, :] = -3.0; lim[1, 7] = 20.0; lim[1, 30] = 30.0                # both saturate to 1.0 -> index 7
    lim[2, :] = -200.0; lim[2, 5] = -150.0                               # every sigmoid is 0 -> index 0
    lim[3, :] = 0.25; lim[3, 50] = np.nan; lim[3, 60] = np.nan           # first NaN
    lim[4, :] = -np.inf; lim[4, 33] = -120.0                             # sigmoid(-120) = 0 = sigmoid(-inf) -> index 0
    lim[5, :] = 0.0; lim[5, 80] = np.inf                                 # +inf -> 1.0, last element
    lim[6, :] = 2.0; lim[6, 3] = up(2.0, 3); lim[6, 9] = up(2.0, 40)     # 40 ulps: a larger sigmoid -> index 9
    lim[7, 0] = np.nan                                                   # NaN first
    bias_t = torch.from_numpy(bias).cuda()
    dec, amax, logits, head = parser.head_gemm_argmax(feat, weight, bias_t, emit=True)
    torch.cuda.synchronize()
    lg = logits.cpu().numpy()
    want = np.broadcast_to(bias[None, :, None, None], lg.shape)
    assert np.array_equal(np.isnan(lg), np.isnan(want))
    assert np.array_equal(bits(lg)[~np.isnan(want)], bits(want)[~np.isnan(want)]), "0 * x + bias must be bias"
    check_against_emitted(g, parser, dec, amax, logits, head)
    a = amax.cpu().numpy()
    s15 = torch.sigmoid(torch.tensor([1.5, float(up(1.5, 1))])).numpy()
    want0 = 10 if s15[0] == s15[1] else 40
    assert (a[:, 0] == want0).all() and (a[:, 1] == 7).all() and (a[:, 2] == 0).all() and (a[:, 3] == 50).all()
    assert (a[:, 4] == 0).all() and (a[:, 5] == 80).all() and (a[:, 7] == 0).all()


def test_head_rejects_bad_arguments():
    from pytorch_pose_proposal_network_b200 import _lib
    g = geometry("cfg1")
    parser = parser_for(g)
    feat, weight, bias = make_layer(g, 1, 48, seed=1)                     # Cin not a multiple of 32
    with pytest.raises(_lib.PPNError):
        parser.head_gemm_argmax(feat, weight, bias)
    feat, weight, bias = make_layer(g, 1, 64, seed=1)
    with pytest.raises(ValueError):
        parser.head_gemm_argmax(feat, weight[:-1].contiguous(), bias)


@pytest.mark.parametrize("operand", ["f16", "bf16"])
@pytest.mark.parametrize("name,B,Cin", [("cfg1", 3, 64), ("cfg1", 5, 512), ("cfg2", 9, 256), ("cfg3", 2, 512), ("native", 2, 512)])
def test_head16_logits_and_epilogue(name, B, Cin, operand):
    """16-bit operand path: packed K-major operands, resident A tile, narrow last channel tile, rows that cross images."""
    g = geometry(name)
    parser = parser_for(g)
    feat, weight, bias = make_layer(g, B, Cin, seed=150 + B)
    dec, amax, logits, head = parser.head_gemm_argmax(feat, weight, bias, emit=True, operand=operand)
    torch.cuda.synchronize()
    ref = conv_fp32(feat, weight, bias)
    err = (logits - ref).abs()
    atol, rtol = TOL16[operand]
    assert bool((err <= atol + rtol * ref.abs()).all()), f"max |logit error| {float(err.max()):.3e} vs the fp32 convolution"
    dt = DT16[operand]
    ref16 = conv_fp32(feat.to(dt).float(), weight.to(dt).float(), bias)
    err = (logits - ref16).abs()
    assert bool((err <= TOL_ACC[0] + TOL_ACC[1] * ref16.abs()).all()), \
        f"max |logit error| {float(err.max()):.3e} vs the convolution of the rounded operands (accumulation only)"
    check_against_emitted(g, parser, dec, amax, logits, head)
    dec2, amax2, _, _ = parser.head_gemm_argmax(feat, weight, bias, emit=False, operand=operand)
    # a channels_last tensor of the operand type is read in place: same operands bit for bit, same result
    feat_cl = feat.to(dt).contiguous(memory_format=torch.channels_last)
    dec3, amax3, logits3, _ = parser.head_gemm_argmax(feat_cl, weight, bias, emit=True, operand=operand)
    torch.cuda.synchronize()
    assert torch.equal(dec2, dec) and torch.equal(amax2.view(torch.int16), amax.view(torch.int16))
    assert torch.equal(dec3, dec) and torch.equal(amax3.view(torch.int16), amax.view(torch.int16))
    assert torch.equal(logits3.view(torch.int32), logits.view(torch.int32))


@pytest.mark.parametrize("operand", ["f16", "bf16"])
@pytest.mark.parametrize("name,B", [("cfg1", 7), ("cfg2", 64), ("cfg3", 8), ("native", 3)])
def test_head16_parse_matches_oracle_on_emitted_head(name, B, operand):
    g = geometry(name)
    parser = parser_for(g)
    feat, weight, bias = make_layer(g, B, 512, seed=177)
    packed, logits, head = parser.parse_features(feat, weight, bias, emit=True, operand=operand)
    torch.cuda.synchronize()
    ref = c_oracle.parse_batch(head.cpu().numpy(), g)
    assert int(ref["counts"][:, 2].sum()) > 0, "degenerate test input: no humans"
    assert_packed_equals_oracle(packed.numpy(), ref, B)
    two_step = parser.parse(head, out=parser.alloc_output(B)).numpy()
    plain = parser.parse_features(feat, weight, bias, out=parser.alloc_output(B), operand=operand).numpy()
    feat_cl = feat.to(DT16[operand]).contiguous(memory_format=torch.channels_last)
    in_place = parser.parse_features(feat_cl, weight, bias, out=parser.alloc_output(B), operand=operand).numpy()
    torch.cuda.synchronize()
    for other in (two_step, plain, in_place):
        assert np.array_equal(other["count"], packed.numpy()["count"])
        assert_packed_equals_oracle(other, ref, B)


def test_head16_many_tiles_persistent_loop():
    """More M tiles than SMs: every CTA loops over several tiles (A-slot and accumulator phases wrap)."""
    g = geometry("cfg2")
    parser = parser_for(g)
    B = 400                                                    # 400 * 144 / 128 = 450 tiles on 148 SMs
    feat, weight, bias = make_layer(g, B, 128, seed=9)
    dec, amax, logits, head = parser.head_gemm_argmax(feat, weight, bias, emit=True, operand="f16")
    torch.cuda.synchronize()
    ref16 = conv_fp32(feat.half().float(), weight.half().float(), bias)
    err = (logits - ref16).abs()
    assert bool((err <= TOL_ACC[0] + TOL_ACC[1] * ref16.abs()).all()), f"max |logit error| {float(err.max()):.3e}"
    dec2, amax2, _, _ = parser.head_gemm_argmax(feat, weight, bias, emit=False, operand="f16")
    torch.cuda.synchronize()
    assert torch.equal(dec2, dec) and torch.equal(amax2.view(torch.int16), amax.view(torch.int16))
    want = torch.argmax(head[:, 6 * g.K:].reshape(B, g.E, g.S, g.H * g.W), dim=2)
    # torch.argmax picks the first maximum like numpy for distinct values; exact ties are checked on a sample by numpy
    hs = head[:16].cpu().numpy()[:, 6 * g.K:].reshape(16, g.E, g.S, g.H * g.W)
    assert np.array_equal(amax2[:16].cpu().numpy().reshape(16, g.E, -1), np.argmax(hs, axis=2).astype(np.uint16))
    mism = (want != amax2.reshape(B, g.E, -1).to(torch.int64)).float().mean()
    assert float(mism) < 1e-4, f"{float(mism):.2e} of the arg-max entries differ from torch.argmax"


def test_head16_rejects_what_it_cannot_take():
    from pytorch_pose_proposal_network_b200 import _lib
    g = geometry("cfg1")
    parser = parser_for(g)
    feat, weight, bias = make_layer(g, 1, 96, seed=1)                     # Cin not a multiple of 64
    with pytest.raises(_lib.PPNError):
        parser.head_gemm_argmax(feat, weight, bias, operand="f16")
    feat, weight, bias = make_layer(g, 1, 64, seed=1)
    with pytest.raises(ValueError):                                       # a bf16 tensor for fp16 operands
        parser.head_gemm_argmax(feat.bfloat16().contiguous(memory_format=torch.channels_last), weight, bias, operand="f16")
    with pytest.raises(ValueError):                                       # 16-bit NCHW is not a layout the kernel reads
        parser.head_gemm_argmax(feat.half(), weight, bias, operand="f16")
    with pytest.raises(ValueError):
        parser.head_gemm_argmax(feat, weight, bias, operand="fp8")


@pytest.mark.parametrize("operand", ["tf32", "f16"])
def test_head_every_epilogue_width_gives_the_same_answer(operand):
    """head.subs = epilogue warps per TMEM lane quadrant: the pieces a window is cut into change with it, the merged
    arg-max map and the decode planes must not (the key maximum is order-independent and exact)."""
    from pytorch_pose_proposal_network_b200 import _lib
    g = geometry("cfg2")
    parser = parser_for(g)
    feat, weight, bias = make_layer(g, 40, 128, seed=23)
    want = None
    try:
        for subs in (1, 2, 4, 6):
            _lib.tune(head_subs=subs)
            dec, amax, _, _ = parser.head_gemm_argmax(feat, weight, bias, operand=operand)
            torch.cuda.synchronize()
            if want is None:
                want = (dec.clone(), amax.clone())
                _, _, logits, head = parser.head_gemm_argmax(feat, weight, bias, emit=True, operand=operand)
                check_against_emitted(g, parser, dec, amax, logits, head)
            else:
                assert torch.equal(dec, want[0]) and torch.equal(amax.view(torch.int16), want[1].view(torch.int16)), f"subs={subs}"
    finally:
        _lib.tune(head_subs=0)


@pytest.mark.parametrize("operand", ["f16", "bf16"])
@pytest.mark.parametrize("name,B,Cin", [("cfg2", 40, 128), ("native", 2, 512)])
def test_head16_four_narrow_accumulators_give_the_same_answer(name, B, Cin, operand):
    """head.acc = 128: four TMEM accumulators of 128 channels (16 KB weight stages) instead of two of 256 — another tiling
    of the same GEMM and another cut of the windows into pieces; decode planes and arg-max map must not change."""
    from pytorch_pose_proposal_network_b200 import _lib
    g = geometry(name)
    parser = parser_for(g)
    feat, weight, bias = make_layer(g, B, Cin, seed=29)
    try:
        dec, amax, _, _ = parser.head_gemm_argmax(feat, weight, bias, operand=operand)
        torch.cuda.synchronize()
        want = (dec.clone(), amax.clone())
        _lib.tune(head_acc=128)
        dec, amax, logits, head = parser.head_gemm_argmax(feat, weight, bias, emit=True, operand=operand)
        torch.cuda.synchronize()
        check_against_emitted(g, parser, dec, amax, logits, head)
        dec, amax, _, _ = parser.head_gemm_argmax(feat, weight, bias, operand=operand)
        torch.cuda.synchronize()
        assert torch.equal(amax.view(torch.int16), want[1].view(torch.int16))
        assert torch.equal(dec, want[0])
    finally:
        _lib.tune(head_acc=256)


@pytest.mark.parametrize("operand", ["tf32", "f16"])
def test_head_small_batch_items_and_graph_replay(operand):
    """One image of the reference's shape: 5 cell tiles, so every tile's channel tiles are spread over the SMs as work
    items (the key maxima merge them); and the whole call replayed from a CUDA graph on a refilled feature buffer."""
    g = geometry("native")
    parser = parser_for(g)
    feat, weight, bias = make_layer(g, 1, 512, seed=31)
    packed, logits, head = parser.parse_features(feat, weight, bias, emit=True, operand=operand)
    torch.cuda.synchronize()
    ref = c_oracle.parse_batch(head.cpu().numpy(), g)
    assert int(ref["counts"][:, 2].sum()) > 0
    assert_packed_equals_oracle(packed.numpy(), ref, 1)
    cap = parser.capture_features(feat, weight, bias, operand=operand)
    feat2, _, _ = make_layer(g, 1, 512, seed=32)
    feat.copy_(feat2)                                                       # the next frame, written in place
    got = cap.replay().numpy()
    torch.cuda.synchronize()
    packed2, _, head2 = parser.parse_features(feat2, weight, bias, emit=True, operand=operand)
    torch.cuda.synchronize()
    ref2 = c_oracle.parse_batch(head2.cpu().numpy(), g)
    assert_packed_equals_oracle(got, ref2, 1)
    assert_packed_equals_oracle(packed2.numpy(), ref2, 1)
