"""Property tests (hypothesis) of the two CPU restatements against each other and against
simple invariants — no GPU, no reference needed."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_oracle, ppn_oracle as O

boxes = st.lists(st.tuples(st.floats(0, 300, width=32), st.floats(0, 300, width=32),
                           st.floats(0, 120, width=32), st.floats(0, 120, width=32)), min_size=0, max_size=60)


@settings(max_examples=150, deadline=None)
@given(boxes, st.sampled_from([0.1, 0.3, 0.5, 0.75]), st.integers(0, 2 ** 31 - 1), st.sampled_from([None, 1, 3, 10]))
def test_nms_numpy_equals_c_and_is_a_valid_greedy_selection(bx, thr, seed, limit):
    b = np.array([[y, x, y + h, x + w] for y, x, h, w in bx], np.float32).reshape(-1, 4)
    n = len(b)
    rng = np.random.default_rng(seed)
    score = rng.permutation(n).astype(np.float32)                       # distinct
    with np.errstate(all="ignore"):
        keep = O.nms(b, thr, score=score, limit=limit)
        assert np.array_equal(keep, c_oracle.nms(b, thr, score=score, limit=limit))
        assert len(set(keep.tolist())) == len(keep)
        assert (np.diff(score[keep]) < 0).all()                          # visiting order = descending score
        if limit is not None:
            assert len(keep) <= limit
        # no kept box suppresses a later kept box; every dropped box (before the limit hit) is suppressed by a kept one
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        for pos, i in enumerate(keep):
            if pos:
                iou = O.iou_one_to_many(b[i], area[i], b[keep[:pos]], area[keep[:pos]])
                assert not (iou >= np.float32(thr)).any()
        if limit is None:
            kept = set(keep.tolist())
            for i in range(n):
                if i not in kept:
                    better = [j for j in keep if score[j] > score[i]]
                    iou = O.iou_one_to_many(b[i], area[i], b[better], area[better])
                    assert (iou >= np.float32(thr)).any()


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 10 ** 6), st.sampled_from(["U", "R", "D", "S"]), st.integers(2, 7), st.integers(2, 7),
       st.sampled_from([1, 3, 5]))
def test_parse_numpy_equals_c(seed, dist, W, H, s):
    from oracle import synth
    graphs = (((0, 1), (1, 2)), ((0, 2), (1, 3)), ((3,), (4,)))
    g = O.Geometry(K=5, E=4, inW=16 * W, inH=16 * H, W=W, H=H, sW=s, sH=s, graphs=graphs)
    head = synth.make_head(g, dist, seed, B=2)
    if not synth.root_scores_distinct(head, g):
        return
    c = c_oracle.parse_batch(head, g)
    for b in range(2):
        p = O.parse_image(head[b], g)
        n = len(p.root_cell)
        assert int(c["counts"][b, 2]) == n
        assert np.array_equal(c["part_cell"][b, :n], p.part_cell)
        assert np.array_equal(c["part_box"][b, :n].view(np.uint32), p.part_box.view(np.uint32))
        # invariants: root is part 0; every present part is above the threshold; root scores descend
        assert (p.part_cell[:, 0] == p.root_cell).all()
        assert (p.part_score[p.part_cell >= 0] >= np.float32(g.det_thresh)).all()
        assert (np.diff(p.part_score[:, 0]) <= 0).all()


def blockwise_nms_model(b, thr, score, limit=None):
    """Host model of the CUDA NMS's control flow (csrc/ppn_kernels.cu, nms_core): boxes in visiting order, 32 at a
    time; a block is resolved from its 32x32 DIAGONAL suppression bits and the `removed` word inherited from earlier
    blocks; only the boxes a block keeps are then tested against the later blocks (a box already removed is skipped).
    The claim checked below: this visits exactly the pairs that matter, i.e. equals the plain greedy loop."""
    n = len(b)
    order = np.lexsort((-np.arange(n), -score.astype(np.float64)))
    sb = b[order]
    area = (sb[:, 2] - sb[:, 0]) * (sb[:, 3] - sb[:, 1])
    thr = np.float32(thr)

    def sup(i, js):                                             # does kept box i suppress the boxes js?
        return O.iou_one_to_many(sb[i], area[i], sb[js], area[js]) >= thr if len(js) else np.zeros(0, bool)

    Wd = (n + 31) // 32
    removed = np.zeros(n, bool)
    keep = []
    for w in range(Wd):
        i0, i1 = 32 * w, min(n, 32 * w + 32)
        diag = {i: sup(i, np.arange(i + 1, i1)) for i in range(i0, i1)}          # built up front for all blocks
        kept_here = []
        for i in range(i0, i1):                                                  # the one-warp scan
            if removed[i]:
                continue
            kept_here.append(i)
            removed[i + 1:i1] |= diag[i]
        if limit is not None and len(keep) + len(kept_here) >= limit:
            keep += kept_here[:limit - len(keep)]
            break
        keep += kept_here
        for i in kept_here:                                                      # survivors against the later blocks
            later = np.arange(i1, n)
            later = later[~removed[later]]
            removed[later] |= sup(i, later)
    return order[np.asarray(keep, np.int64)].astype(np.int32) if keep else np.zeros(0, np.int32)


wide_boxes = st.lists(st.tuples(st.floats(0, 300, width=32), st.floats(0, 300, width=32),
                                st.floats(0, 200, width=32), st.floats(0, 200, width=32)), min_size=0, max_size=150)


@settings(max_examples=120, deadline=None)
@given(wide_boxes, st.sampled_from([0.05, 0.3, 0.5, 0.9]), st.integers(0, 2 ** 31 - 1), st.sampled_from([None, 2, 40]),
       st.booleans())
def test_blockwise_nms_equals_greedy(bx, thr, seed, limit, poison):
    b = np.array([[y, x, y + h, x + w] for y, x, h, w in bx], np.float32).reshape(-1, 4)
    n = len(b)
    rng = np.random.default_rng(seed)
    if poison and n:                                            # zero-area, inverted and NaN boxes (NaN IoU keeps a box)
        k = rng.integers(0, n, size=max(1, n // 6))
        b[k, 2] = b[k, 0]
        b[k[: len(k) // 2], 3] = np.nan
    score = rng.permutation(n).astype(np.float32)
    with np.errstate(all="ignore"):
        assert np.array_equal(blockwise_nms_model(b, thr, score, limit), O.nms(b, thr, score=score, limit=limit))


def matrix_nms_model(b, thr, score, limit=None):
    """Host model of an NMS formulation the CUDA path tried and does NOT use (nms_core's comment): the
    full suppression matrix first — row i, 32-bit words, bit j set when box i suppresses the LATER box j, words before
    i's own never written — then a walk over the survivors: the first box not in the removed set is kept, marked
    visited, and the words of its row from its own word on are OR-ed into the set.  Exact, but slower on the GPU than
    the block-by-block scheme above; kept as a property check of the formulation."""
    n = len(b)
    order = np.lexsort((-np.arange(n), -score.astype(np.float64)))
    sb = b[order]
    area = (sb[:, 2] - sb[:, 0]) * (sb[:, 3] - sb[:, 1])
    thr = np.float32(thr)
    Wd = (n + 31) // 32
    rng = np.random.default_rng(n)
    mat = rng.integers(0, 2 ** 32, size=(n, max(Wd, 1)), dtype=np.uint64)          # unwritten words hold garbage
    for i in range(n):
        later = np.arange(i + 1, n)
        bits = np.zeros(Wd * 32, bool)
        if len(later):
            bits[later] = O.iou_one_to_many(sb[i], area[i], sb[later], area[later]) >= thr
        words = np.packbits(bits.reshape(Wd, 32), axis=1, bitorder="little").view("<u4").ravel().astype(np.uint64)
        mat[i, i // 32:] = words[i // 32:]
    rem = np.array([0 if n - 32 * w >= 32 else (0xffffffff & ~((1 << max(n - 32 * w, 0)) - 1)) for w in range(Wd)], np.uint64)
    keep = []
    while True:
        free = [w for w in range(Wd) if rem[w] != 0xffffffff]
        if not free:
            break
        fw = free[0]
        fb = (~int(rem[fw])) & 0xffffffff
        i = 32 * fw + (fb & -fb).bit_length() - 1
        keep.append(i)
        if limit is not None and len(keep) >= limit:
            break
        rem[fw:] |= mat[i, fw:]
        rem[fw] |= np.uint64(1 << (i & 31))
    return order[np.asarray(keep, np.int64)].astype(np.int32) if keep else np.zeros(0, np.int32)


@settings(max_examples=120, deadline=None)
@given(wide_boxes, st.sampled_from([0.05, 0.3, 0.5, 0.9]), st.integers(0, 2 ** 31 - 1), st.sampled_from([None, 2, 40]),
       st.booleans())
def test_matrix_nms_equals_greedy(bx, thr, seed, limit, poison):
    b = np.array([[y, x, y + h, x + w] for y, x, h, w in bx], np.float32).reshape(-1, 4)
    n = len(b)
    rng = np.random.default_rng(seed)
    if poison and n:
        k = rng.integers(0, n, size=max(1, n // 6))
        b[k, 2] = b[k, 0]
        b[k[: len(k) // 2], 3] = np.nan
    score = rng.permutation(n).astype(np.float32)
    with np.errstate(all="ignore"):
        assert np.array_equal(matrix_nms_model(b, thr, score, limit), O.nms(b, thr, score=score, limit=limit))


def wavefront_nms_model(b, thr, score, limit=None):
    """Host model of the control flow of nms_core's default path (csrc/ppn_kernels.cu): ranks from a 128-bucket sort on
    the upper key word (min/max, shift, counts, suffix sums, the n x n count only inside a bucket), the boxes visited a
    32-box word at a time; a word tests its boxes against the survivor LIST of the earlier words only, then settles
    its own dependency from the TRANSPOSED diagonal bits (column t = the earlier boxes of the word that would suppress
    box t) by rounds: a box with a kept suppressor dies, a box none of whose suppressors is still undecided is kept."""
    n = len(b)
    if n == 0:
        return np.zeros(0, np.int32)
    # keys as the kernels build them: monotone image of the fp32 score in the upper word, the index in the lower
    u = score.astype(np.float32).view(np.uint32).astype(np.uint64)
    hi = np.where(u & 0x80000000, ~u & 0xffffffff, u | 0x80000000).astype(np.uint64)
    key = (hi << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    lo_, hi_ = int(hi.min()), int(hi.max())
    rng_ = hi_ - lo_
    clz = 32 - rng_.bit_length()
    shift = max(0, 25 - clz)
    bucket = ((hi.astype(np.int64) - lo_) >> shift).astype(np.int64)
    assert bucket.max() < 128
    counts = np.bincount(bucket, minlength=128)
    above = np.concatenate([np.cumsum(counts[::-1])[::-1][1:], [0]])           # boxes in higher buckets
    rank = np.empty(n, np.int64)
    for i in range(n):
        same = np.nonzero(bucket == bucket[i])[0]
        rank[i] = above[bucket[i]] + int((key[same] > key[i]).sum())
    assert sorted(rank.tolist()) == list(range(n))                             # a permutation
    order = np.empty(n, np.int64)
    order[rank] = np.arange(n)
    sb = b[order]
    area = (sb[:, 2] - sb[:, 0]) * (sb[:, 3] - sb[:, 1])
    thr = np.float32(thr)
    survivors = []                                                             # positions in visiting order
    for i0 in range(0, n, 32):
        idx = np.arange(i0, min(n, i0 + 32))
        nb = len(idx)
        dead = np.zeros(nb, bool)
        for q in survivors:                                                    # the list published by the earlier words
            dead |= O.iou_one_to_many(sb[q], area[q], sb[idx], area[idx]) >= thr
        row = np.zeros((nb, nb), bool)                                         # row[i, t]: box i suppresses the later box t
        for i in range(nb):
            if i + 1 < nb:
                row[i, i + 1:] = O.iou_one_to_many(sb[idx[i]], area[idx[i]], sb[idx[i + 1:]], area[idx[i + 1:]]) >= thr
        col = row.T                                                            # col[t, i]: the transposed bits a lane holds
        und, kept = ~dead, np.zeros(nb, bool)
        rounds = 0
        while und.any():
            killed = und & (col & kept[None, :]).any(axis=1)
            free = und & ~killed & ~(col & und[None, :]).any(axis=1)
            assert (killed | free).any()                                       # the first undecided box always decides
            kept |= free
            und &= ~(killed | free)
            rounds += 1
        assert rounds <= 32
        new = idx[kept].tolist()
        if limit is not None and len(survivors) + len(new) >= limit:
            survivors += new[:max(limit - len(survivors), 0)]
            break
        survivors += new
    return order[np.asarray(survivors, np.int64)].astype(np.int32) if survivors else np.zeros(0, np.int32)


@settings(max_examples=120, deadline=None)
@given(wide_boxes, st.sampled_from([0.05, 0.3, 0.5, 0.9]), st.integers(0, 2 ** 31 - 1), st.sampled_from([None, 1, 2, 40]),
       st.booleans(), st.sampled_from(["perm", "few", "equal", "uniform"]))
def test_wavefront_nms_equals_greedy(bx, thr, seed, limit, poison, scores):
    b = np.array([[y, x, y + h, x + w] for y, x, h, w in bx], np.float32).reshape(-1, 4)
    n = len(b)
    rng = np.random.default_rng(seed)
    if poison and n:
        k = rng.integers(0, n, size=max(1, n // 6))
        b[k, 2] = b[k, 0]
        b[k[: len(k) // 2], 3] = np.nan
    score = {"perm": lambda: rng.permutation(n).astype(np.float32),
             "few": lambda: rng.integers(0, 3, n).astype(np.float32),            # many ties: long buckets, index order decides
             "equal": lambda: np.full(n, 0.5, np.float32),                       # one bucket
             "uniform": lambda: (0.15 + 0.85 * rng.random(n)).astype(np.float32)}[scores]()
    with np.errstate(all="ignore"):
        got = wavefront_nms_model(b, thr, score, limit)
        # the reference's order of exactly equal scores is unspecified (argsort); the rule here is larger index first
        order = np.lexsort((-np.arange(n), -score.astype(np.float64)))
        want = order[O.nms(b[order], thr, score=None, limit=limit)] if n else np.zeros(0, np.int32)
        assert np.array_equal(got, want.astype(np.int32))
