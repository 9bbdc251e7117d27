"""CPU-only tests: configuration, the C-ABI library's symbols and argument checks (no compute
calls), sharding arithmetic and the world_size-2 gather over gloo."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_presets_match_baseline_sizes():
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    expect = {"mpii16": (1311, 755136), "coco18": (1485, 1520640), "highres": (2165, 4988160),
              "reference_native": (7605, 17521920)}
    for name, (Cn, nbytes) in expect.items():
        cfg = getattr(PPNConfig, name)()
        assert cfg.C == Cn and cfg.bytes_per_image == nbytes, name
    cfg = PPNConfig.reference_native()
    assert cfg.gridsize == (16, 16) and cfg.off_h == 10 and cfg.off_w == 10
    assert cfg.key_order() == [0, 15, 13, 1, 3, 5, 2, 4, 6, 17, 14, 7, 9, 11, 8, 10, 12, 16]


def test_skeleton_values():
    from pytorch_pose_proposal_network_b200 import config as c
    assert len(c.KEYPOINT_NAMES) == 18 and len(c.EDGES) == 17
    assert c.EDGES == [[0, 15], [15, 13], [13, 1], [1, 3], [3, 5], [13, 2], [2, 4], [4, 6], [13, 17], [17, 14],
                       [14, 7], [14, 8], [7, 9], [8, 10], [9, 11], [10, 12], [0, 16]]
    assert c.DIRECTED_GRAPHS == [[[0, 1, 2, 3, 4], [15, 13, 1, 3, 5]], [[0, 1, 5, 6, 7], [15, 13, 2, 4, 6]],
                                 [[0, 1, 8, 9, 10, 12, 14], [15, 13, 17, 14, 7, 9, 11]],
                                 [[0, 1, 8, 9, 11, 13, 15], [15, 13, 17, 14, 8, 10, 12]], [[16], [16]]]
    assert len(c.KEYPOINT_NAMES_16) == 16 and len(c.EDGES_16) == 15
    # every 16-part limb is reached by exactly one track-order step sequence (a tree)
    reached = {t for _, ts in c.DIRECTED_GRAPHS_16 for t in ts}
    assert reached == set(range(1, 16))


def test_pairwise():
    from pytorch_pose_proposal_network_b200.utils import pairwise
    assert list(pairwise("abcd")) == [("a", "b"), ("b", "c"), ("c", "d")]
    assert list(pairwise([1])) == [] and list(pairwise([])) == []


def test_config_validation():
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    with pytest.raises(ValueError):
        PPNConfig(K=3, E=1, directed_graphs=(((0,), (5,)),))
    with pytest.raises(ValueError):
        PPNConfig(K=3, E=1, directed_graphs=(((0, 0), (1,)),))
    off, limb, part = PPNConfig.reference_native().chains()
    assert list(off) == [0, 5, 10, 17, 24, 25] and len(limb) == 25 and len(part) == 25


def test_color_map_is_the_references():
    """config.py:23-42 — digest of the reference's own COLOR_MAP (oracle/make_golden.py asserts equality live)."""
    import hashlib
    from pytorch_pose_proposal_network_b200 import config as c
    assert list(c.COLOR_MAP)[:4] == ["instance", "right_shoulder", "right_elbow", "right_wrist"]
    assert set(c.COLOR_MAP) == set(c.KEYPOINT_NAMES) and c.COLOR_MAP["instance"] == (143, 35, 35)
    assert hashlib.sha256(repr(list(c.COLOR_MAP.items())).encode()).hexdigest() == \
        "f5bd531494d0da561eb0b00dbbfb89276dacc417e26f61cf92804303fad6065c"


def test_library_exports_every_declared_symbol():
    from pytorch_pose_proposal_network_b200 import _lib
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("ppn_decode.h", "ppn_decode_bench.h"))
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    product = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ppn_decode.h")).read(), flags=re.S)
    assert "ppn_tune" not in product and "ppn_profile" not in product, "benchmark hooks belong in ppn_decode_bench.h"
    declared = set(re.findall(r"\b(ppn_[a-z_0-9]+)\s*\(", body))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.ppn_abi_version() == _lib.ABI_VERSION == 7
    assert b"workspace" in lib.ppn_strerror(-3)


def test_abi_argument_checks_without_gpu():
    """Pure host-side validation paths: no kernel is launched."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import _CConfig
    lib = _lib.lib()
    cc = _CConfig(PPNConfig.mpii16())
    shape = cc.shape(512)
    need = C.c_size_t()
    assert lib.ppn_workspace_bytes(C.byref(shape), C.byref(cc.params), C.byref(need)) == 0
    # two sets (used alternately) of: arg-max map + surviving root cells + counts
    B, HW, E = 512, 144, 15
    one = B * E * HW * 2 + B * HW * 4 + B * 4
    assert 2 * one <= need.value <= 2 * (one + 3 * 256)
    assert lib.ppn_workspace_bytes(None, C.byref(cc.params), C.byref(need)) == -1
    bad = cc.shape(1); bad.K = 0
    assert lib.ppn_workspace_bytes(C.byref(bad), C.byref(cc.params), C.byref(need)) == -1
    big = cc.shape(1); big.sH = big.sW = 300
    assert lib.ppn_workspace_bytes(C.byref(big), C.byref(cc.params), C.byref(need)) == -2
    hs = _lib.PPNHumans(0, 0, 0, 0, 0, 4)
    assert lib.ppn_parse(None, C.byref(shape), C.byref(cc.params), C.byref(hs), None, 0, None) == -1
    assert lib.ppn_parse_launches(C.byref(shape), C.byref(cc.params)) >= 1
    assert lib.ppn_tune(b"no.such.knob", 1) == -1
    v = C.c_int32()
    assert lib.ppn_tune_get(b"argmax.stages", C.byref(v)) == 0 and v.value >= 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_silent_cpu_fallback():
    from pytorch_pose_proposal_network_b200 import datatest as dt
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser
    with pytest.raises(RuntimeError):
        PoseParser(PPNConfig.mpii16())
    with pytest.raises(RuntimeError):
        dt.non_maximum_suppression(np.zeros((2, 4), np.float32), 0.3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pytorch_pose_proposal_network_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("test oracle", ""), os.path.join(dirpath, f)


def test_shard_ranges_cover_job():
    from pytorch_pose_proposal_network_b200.sharded import shard_range, shard_sizes
    for n in (0, 1, 7, 512, 8192, 8191):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert sum(shard_sizes(n, world)) == n
            assert max(shard_sizes(n, world)) == -(-n // world)


def test_dense_entry_layout_host_side():
    """unpack_entries / entries_to_packed against a buffer laid out by hand from ppn_packed_bytes' offsets (no GPU)."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.parser import entries_to_packed, unpack_entries
    B, K, cap = 3, 4, 16
    nbytes, offs = C.c_size_t(), (C.c_size_t * 4)()
    assert _lib.lib().ppn_packed_bytes(B, cap, C.byref(nbytes), offs) == 0
    offs = tuple(int(o) for o in offs)
    assert offs[0] == 0 and all(o % 256 == 0 for o in offs) and offs[1] >= 4 * (2 + 3 * B)
    assert nbytes.value >= offs[3] + cap * 16
    buf = np.zeros(nbytes.value, np.uint8)
    # image 0: humans {0,2} and {0}; image 1: none; image 2: human {0,1,3}
    parts = [0, 2, 0, 0, 1, 3]
    cells = [5, 6, 9, 1, 2, 3]
    count, entries = np.array([2, 0, 1], np.int32), np.array([3, 0, 3], np.int32)
    buf[0:4 * (2 + 3 * B)].view(np.int32)[:] = np.concatenate([[6, 0], count, entries, [0, 3, 3]])
    buf[offs[1]:offs[1] + 4 * 6].view(np.uint32)[:] = [(p << 16) | c for p, c in zip(parts, cells)]
    buf[offs[2]:offs[2] + 4 * 6].view(np.float32)[:] = np.arange(6, dtype=np.float32) / 8
    rec = unpack_entries(buf, B, cap, offs)
    assert rec["total"] == 6 and not rec["overflow"] and list(rec["start"]) == [0, 3, 3]
    pc, ps, pb = entries_to_packed(rec, 0, K)
    assert pc.tolist() == [[5, -1, 6, -1], [9, -1, -1, -1]] and ps[0, 2] == np.float32(1 / 8)
    assert entries_to_packed(rec, 1, K)[0].shape == (0, K)
    assert entries_to_packed(rec, 2, K)[0].tolist() == [[1, 2, -1, 3]]
    assert _lib.lib().ppn_packed_bytes(-1, cap, C.byref(nbytes), None) == -1


def _gloo_worker(rank, world, port, n_images, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import c_oracle, ppn_oracle as O, synth
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PackedHumans
    from pytorch_pose_proposal_network_b200.sharded import gather_packed, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = PPNConfig.mpii16()
        g = O.Geometry.of(cfg)
        head = synth.make_head(g, "U", seed=77, B=n_images)          # same job on every rank
        lo, hi = shard_range(n_images, world, rank)
        # stand-in for the GPU parser on this rank's block: the oracle (tests may use it)
        r = c_oracle.parse_batch(head[lo:hi], g) if hi > lo else None
        K, HW = cfg.K, cfg.HW
        def t(a, shape, dtype): return torch.from_numpy(a) if a is not None else torch.zeros(shape, dtype=dtype)
        local = PackedHumans(cfg,
                             t(r["counts"][:, 2].copy() if r else None, (0,), torch.int32),
                             t(r["root_cell"] if r else None, (0, HW), torch.int32),
                             t(r["part_cell"] if r else None, (0, HW, K), torch.int32),
                             t(r["part_score"] if r else None, (0, HW, K), torch.float32),
                             t(r["part_box"] if r else None, (0, HW, K, 4), torch.float32))
        full = gather_packed(local, n_images, trim_humans=64)
        whole = c_oracle.parse_batch(head, g)
        got = full.numpy()
        ok = np.array_equal(got["count"], whole["counts"][:, 2])
        for b in range(n_images):
            n = int(whole["counts"][b, 2])
            ok &= np.array_equal(got["part_cell"][b, :n], whole["part_cell"][b, :n])
            ok &= np.array_equal(got["part_box"][b, :n].view(np.uint32), whole["part_box"][b, :n].view(np.uint32))
        open(os.path.join(tmp, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [10, 7])
def test_gather_world2_gloo(tmp_path, n_images):
    """Sharded == unsharded, through the real all_gather with two processes (gloo, CPU)."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + n_images
    mp.spawn(_gloo_worker, args=(2, port, n_images, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]


def test_argmax_work_items_tile_the_batch():
    """The ring kernel's work items (full items, then a tail of quarter items) must cover every (image, limb)
    matrix exactly once, in order — pure host arithmetic of the launch plan, checked without a GPU."""
    from pytorch_pose_proposal_network_b200 import _lib
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import _CConfig
    lib = _lib.lib()
    for cfg, batches in ((PPNConfig.mpii16(), (1, 7, 64, 300, 512, 513, 2048)), (PPNConfig.coco18(), (3, 1024)),
                         (PPNConfig.highres(), (256,)), (PPNConfig.reference_native(), (1, 64, 65))):
        cc = _CConfig(cfg)
        for B in batches:
            for sms in (148, 132, 7):
                shape = cc.shape(B)
                info = (C.c_int32 * 4)()
                n_mats = B * cfg.E
                first, size = (C.c_int32 * (n_mats + 1))(), (C.c_int32 * (n_mats + 1))()
                rc = lib.ppn_debug_argmax_items(C.byref(shape), sms, info, first, size, n_mats + 1)
                if rc == -2:
                    continue                    # this shape does not take the split-matrix ring
                assert rc == 0
                G, n_big, small_m, n_items = list(info)
                assert 1 <= small_m <= G and 0 <= n_big <= n_items <= n_mats
                at = 0
                for it in range(n_items):
                    assert first[it] == at and 1 <= size[it] <= (G if it < n_big else small_m), (B, sms, it)
                    at += size[it]
                assert at == n_mats, (B, sms, at, n_mats)
