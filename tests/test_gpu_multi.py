"""Image-sharded parsing on 2 GPUs with the GPU parser on every rank (BASELINE.json configs[4]):
the poses gathered at rank 0 — by kernel stores into its peer-mapped buffer, by copy-engine copies, or by NCCL
all_gather — must equal, bit for bit, what ONE GPU produces for the same images.

Needs two visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped otherwise.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CHUNK, CHUNKS_PER_RANK, WORLD = 48, 5, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _chunk(cfg, c, dev):
    gen = torch.Generator(device=dev).manual_seed(31000 + c)
    return torch.rand(CHUNK, cfg.C, cfg.H, cfg.W, device=dev, generator=gen)


def _worker(rank, port, mode, result_path):
    import torch.distributed as dist
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser, entries_to_packed
    from pytorch_pose_proposal_network_b200.sharded import PeerPoseGatherer, PoseGatherer, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    try:
        cfg = PPNConfig.mpii16()
        parser = PoseParser(cfg, device=dev)
        n_images = CHUNK * CHUNKS_PER_RANK * WORLD
        lo, hi = shard_range(n_images, WORLD, rank)
        mine = list(range(lo // CHUNK, hi // CHUNK))
        cap = CHUNK * 6 * cfg.K
        if mode == "nccl":
            g = PoseGatherer(parser, CHUNK, cap, group_steps=2)            # 5 chunks: two full groups and a flushed one
        else:                                                              # "store" / "copy" [+ "-nccl": the all_gather notification]
            g = PeerPoseGatherer(parser, CHUNK, cap, slots=8, notify_every=2, mode=mode.split("-")[0],
                                 control="nccl" if mode.endswith("-nccl") else "flags")
        outs = [parser.alloc_output(CHUNK) for _ in range(2)]
        heads = [_chunk(cfg, c, dev) for c in mine]
        torch.cuda.synchronize(dev)
        for i, h in enumerate(heads):
            g.parse(h, out=outs[i % 2], input_complete=True)
        g.finish()
        torch.cuda.synchronize(dev)
        if hasattr(g, "check_landed"):
            g.check_landed()
        dist.barrier()
        ok, detail = True, ""
        if rank == 0:
            for r in range(WORLD):
                for i in range(CHUNKS_PER_RANK):
                    c = r * CHUNKS_PER_RANK + i
                    if mode == "nccl" and i < 2:
                        continue                                           # the all_gather variant holds the two most recent groups only
                    rec = g.records_of(r, step_back=CHUNKS_PER_RANK - 1 - i)
                    assert not rec["overflow"]
                    ref = parser.parse(_chunk(cfg, c, dev), out=parser.alloc_output(CHUNK)).numpy()   # ONE GPU, same images
                    torch.cuda.synchronize(dev)
                    if not np.array_equal(rec["count"], ref["count"]):
                        ok, detail = False, f"counts of chunk {c} differ"
                        break
                    for b in range(CHUNK):
                        n = int(ref["count"][b])
                        pc, ps, pb = entries_to_packed(rec, b, cfg.K)
                        same = (pc.shape[0] == n and np.array_equal(pc, ref["part_cell"][b, :n])
                                and np.array_equal(ps.view(np.uint32), ref["part_score"][b, :n].view(np.uint32))
                                and np.array_equal(pb.view(np.uint32), ref["part_box"][b, :n].view(np.uint32)))
                        if not same:
                            ok, detail = False, f"image {b} of chunk {c} (rank {r}) differs"
                            break
                    if not ok:
                        break
                if not ok:
                    break
            with open(result_path, "w") as f:
                f.write("ok" if ok else "FAIL: " + detail)
        dist.barrier()
        if hasattr(g, "close"):
            g.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["store", "copy", "nccl", "store-nccl", "copy-nccl"])
def test_two_gpus_gathered_equals_single_gpu(mode, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    result = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(_free_port(), mode, result), nprocs=WORLD, join=True)
    assert open(result).read() == "ok"


def test_remote_dense_records_on_one_gpu():
    """ppn_parse_dense_remote with the 'remote' buffer in local memory: same records as ppn_parse_dense."""
    from pytorch_pose_proposal_network_b200.config import PPNConfig
    from pytorch_pose_proposal_network_b200.parser import PoseParser, entries_to_packed, unpack_entries
    cfg = PPNConfig.mpii16()
    parser = PoseParser(cfg)
    B, cap = 40, 40 * 6 * cfg.K
    head = _chunk(cfg, 3, parser.device)[:B].contiguous()
    nbytes, offs = parser.packed_layout(B, cap)
    local = torch.zeros(nbytes, dtype=torch.uint8, device=parser.device)
    remote = torch.zeros(nbytes, dtype=torch.uint8, device=parser.device)
    plain = torch.zeros(nbytes, dtype=torch.uint8, device=parser.device)
    out = parser.parse(head, out=parser.alloc_output(B), dense=local, cap_entries=cap, remote=(remote.data_ptr(), nbytes))
    ref_out = parser.parse(head, out=parser.alloc_output(B), dense=plain, cap_entries=cap)
    torch.cuda.synchronize()
    rr = unpack_entries(remote.cpu(), B, cap, offs, derive=True)
    pp = unpack_entries(plain.cpu(), B, cap, offs)
    assert rr["total"] == pp["total"] and not rr["overflow"] and np.array_equal(rr["count"], pp["count"])
    assert np.array_equal(out.count.cpu().numpy(), ref_out.count.cpu().numpy())
    for b in range(B):
        for a, c in zip(entries_to_packed(rr, b, cfg.K), entries_to_packed(pp, b, cfg.K)):
            assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, c.view(np.uint32) if c.dtype == np.float32 else c)
    # an entry buffer that is too small: the images that do not fit are reported, nothing is written out of bounds
    small_cap = max(1, pp["total"] // 3)
    nb2, offs2 = parser.packed_layout(B, small_cap)
    remote2 = torch.zeros(nb2 + 4096, dtype=torch.uint8, device=parser.device)
    parser.parse(head, out=parser.alloc_output(B), dense=local, cap_entries=small_cap, remote=(remote2.data_ptr(), nb2))
    torch.cuda.synchronize()
    assert unpack_entries(remote2[:nb2].cpu(), B, small_cap, offs2, derive=True)["overflow"]
    assert int(remote2[nb2:].sum()) == 0
