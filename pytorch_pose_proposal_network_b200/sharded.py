"""Image-sharded multi-GPU parsing: one process per GPU, no data-path collective.

Images are independent (the reference parses them one at a time, rt_test.py:133), so a job of
N images is cut into contiguous blocks, rank r owning images ``[r*n, (r+1)*n)`` with
``n = ceil(N / world)``; every rank runs the single-GPU pipeline on its block.  The only
communication is the gather of the per-rank pose lists (and of timings in bench.py): an
``all_gather`` of the per-image counts and of the fixed-stride packed records, over NCCL on
NVLink/NVSwitch on a GPU box and over gloo in the CPU tests.

The reference's own use of torch.distributed (main.py:240-245, 1233-1238) is data-parallel
training; nothing of it is on this path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .config import PPNConfig


def shard_range(n_images: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of images owned by ``rank``: [lo, hi)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    per = -(-n_images // world)
    lo = min(rank * per, n_images)
    return lo, min(lo + per, n_images)


def shard_sizes(n_images: int, world: int):
    return [hi - lo for lo, hi in (shard_range(n_images, world, r) for r in range(world))]


_FIELDS = ("count", "root_cell", "part_cell", "part_score", "part_box")


def gather_packed(local, n_images: int, group=None, trim_humans: Optional[int] = None):
    """All-gather the packed humans of every rank's shard into the full job's result.

    ``local`` is a :class:`..parser.PackedHumans` for this rank's block (device tensors with
    NCCL, CPU tensors with gloo).  Every rank returns the same full-size ``PackedHumans``; images
    keep their global order because blocks are contiguous.  ``trim_humans`` gathers only the
    first that many slots per image (callers that know an upper bound on humans per image cut
    the payload with it; ``count`` still reports the true number).
    """
    from .parser import PackedHumans
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_images, world)
    per = max(sizes) if sizes else 0
    lo, hi = shard_range(n_images, world, rank)
    if local.count.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.count.shape[0]} images, its shard is {hi - lo}")
    R = local.R if trim_humans is None else min(local.R, int(trim_humans))
    gathered = {}
    for name in _FIELDS:
        t = getattr(local, name)
        if name != "count":
            t = t[:, :R]
        if t.shape[0] < per:                                   # last rank(s): pad to the common block size
            pad = torch.zeros((per - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat([t, pad], dim=0)
        t = t.contiguous()
        full = torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t, group=group)
        if world * per != n_images:
            full = torch.cat([full[r * per: r * per + sizes[r]] for r in range(world)], dim=0)
        gathered[name] = full
    return PackedHumans(local.cfg, *(gathered[n] for n in _FIELDS))


class ShardedPoseParser:
    """Streams this rank's block of a job through a :class:`PoseParser` in chunks."""

    def __init__(self, cfg: PPNConfig, device=None, chunk_images: int = 512, max_humans: Optional[int] = None):
        from .parser import PoseParser
        self.cfg = cfg
        self.chunk = int(chunk_images)
        self.parser = PoseParser(cfg, device=device, max_humans=max_humans)

    def parse_block(self, head: torch.Tensor, out=None):
        """head: this rank's images [n, C, H, W] on the device.  Chunks run back to back on the
        current stream, each writing its slice of one output block (no host sync in between)."""
        from .parser import PackedHumans
        n = head.shape[0]
        if out is None:
            out = self.parser.alloc_output(n)
        for b0 in range(0, n, self.chunk):
            b1 = min(n, b0 + self.chunk)
            view = PackedHumans(self.cfg, out.count[b0:b1], out.root_cell[b0:b1], out.part_cell[b0:b1],
                                out.part_score[b0:b1], out.part_box[b0:b1])
            self.parser.parse(head[b0:b1], out=view)
        return out


class PoseGatherer:
    """Gather of every rank's poses as ONE small asynchronous collective per group of steps.

    Each step's parse writes its poses directly as dense records into a slice of a group buffer
    (``PoseParser.parse(dense=...)`` -> ``ppn_parse_dense``: counts + one (part, cell, score, box) entry per
    present part, up to ``cap_entries``) — by the parse kernel itself, so a step puts nothing but its
    two kernels on the compute stream and consecutive steps stay overlapped.  Every ``group_steps``
    steps the group's buffer is all-gathered with a single ``all_gather_into_tensor`` issued with
    ``async_op`` on a side stream (one event per GROUP, not per step): the collective runs while the next
    steps' kernels run on the compute stream (one NCCL call costs tens of µs of host time, comparable to
    a whole step, hence the grouping).  Two buffer sets alternate; before a set is reused the compute
    stream waits for the collective that last read it.  Every rank ends up with every rank's records of
    every step.
    """

    def __init__(self, parser, images_per_rank: int, cap_entries: int, group=None, group_steps: int = 1):
        self.parser = parser
        self.group = group
        self.world = dist.get_world_size(group)
        self.B = int(images_per_rank)
        self.cap = int(cap_entries)
        self.gs = max(1, int(group_steps))
        self.nbytes, self.offsets = parser.packed_layout(self.B, self.cap)
        dev = parser.device
        self.local = [torch.zeros(self.gs * self.nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.full = [torch.zeros(self.world * self.gs * self.nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.work = [None, None]
        self.step = 0                             # slot cursor: finish() rounds it up to a whole group
        self._last = -1                           # slot index of the last step really submitted
        self.side = torch.cuda.Stream(device=dev)
        self._filled = [torch.cuda.Event() for _ in range(2)]
        self._parsed = [torch.cuda.Event() for _ in range(4)]
        self._released = [torch.cuda.Event() for _ in range(4)]
        self._slices = [[self.local[i][k * self.nbytes:(k + 1) * self.nbytes] for k in range(self.gs)] for i in range(2)]
        self._shipped = [self.gs, self.gs]        # slices per rank in the last gather of each buffer set

    def _ship(self, i: int, n_slices: Optional[int] = None):
        """All-gather buffer set i on the side stream once the compute stream has filled it.  A partly
        filled group (flush) ships only its used slices: rank r's slice k then sits at (r * n + k)."""
        n = self.gs if n_slices is None else int(n_slices)
        self._shipped[i] = n
        main = torch.cuda.current_stream(self.parser.device)
        self._filled[i].record(main)
        self.side.wait_event(self._filled[i])
        with torch.cuda.stream(self.side):
            self.work[i] = dist.all_gather_into_tensor(self.full[i][:self.world * n * self.nbytes],
                                                       self.local[i][:n * self.nbytes], group=self.group, async_op=True)

    def parse(self, head, out=None, input_complete: bool = False):
        """The step: parse `head` on the current stream with the poses written straight into this step's
        slice of the group buffer; when a group is full, start its gather.  Returns the PackedHumans whose
        ``count`` is valid (the fixed-stride arrays are skipped where the kernel can write the dense
        records itself)."""
        grp, k = divmod(self.step, self.gs)
        i = grp & 1
        if k == 0 and self.work[i] is not None:
            self.work[i].wait()                   # compute stream waits for the gather that last read this set
            self.work[i] = None
        res = self.parser.parse(head, out=out, input_complete=input_complete, dense=self._slices[i][k],
                                cap_entries=self.cap, skip_slots=True)
        if k == self.gs - 1:
            self._ship(i)
        self._last = self.step
        self.step += 1
        return res

    def submit(self, humans) -> torch.cuda.Event:
        """Alternative to :meth:`parse` for a result that already exists: pack `humans` (this rank's
        PackedHumans of the step) on the side stream and, when a group is full, start its gather.
        Returns an event that fires once `humans` has been read: wait for it (``stream.wait_event``)
        before the parser overwrites that output buffer.  (Per-step events on the compute stream: steps
        no longer overlap one another.)"""
        grp, k = divmod(self.step, self.gs)
        i = grp & 1
        e = self.step & 3
        parsed, released = self._parsed[e], self._released[e]
        parsed.record(torch.cuda.current_stream(self.parser.device))
        self.side.wait_event(parsed)
        if k == 0 and self.work[i] is not None:
            with torch.cuda.stream(self.side):
                self.work[i].wait()               # side stream waits for the gather that last read this set
            self.work[i] = None
        self.parser.pack(humans, self.cap, buf=self._slices[i][k], stream=self.side)
        released.record(self.side)
        if k == self.gs - 1:
            with torch.cuda.stream(self.side):
                self._shipped[i] = self.gs
                self.work[i] = dist.all_gather_into_tensor(self.full[i], self.local[i], group=self.group, async_op=True)
        self._last = self.step
        self.step += 1
        return released

    def finish(self):
        """Gather a partly filled last group, then make the CURRENT stream wait for every gather.
        Every rank must have submitted the same number of steps."""
        grp, k = divmod(self.step, self.gs)
        if k != 0:                                # flush: ship the slices of the partial group that were filled
            self._ship(grp & 1, n_slices=k)
            self.step = (grp + 1) * self.gs
        with torch.cuda.stream(self.side):
            for j in range(2):
                if self.work[j] is not None:
                    self.work[j].wait()
                    self.work[j] = None
            landed = torch.cuda.Event()
            landed.record(self.side)
        torch.cuda.current_stream(self.parser.device).wait_event(landed)

    def records_of(self, rank: int, step_back: int = 0):
        """Host view of `rank`'s records for the last submitted step minus `step_back`
        (within the two most recent groups; call after finish(); synchronises)."""
        from .parser import unpack_entries
        if self._last < 0 or step_back < 0 or step_back > self._last:
            raise ValueError("no such step")
        grp, k = divmod(self._last - step_back, self.gs)
        if grp < self._last // self.gs - 1:
            raise ValueError("only the two most recent groups of steps are still held")
        buf = self.full[grp & 1]
        lo = (rank * self._shipped[grp & 1] + k) * self.nbytes
        part = buf[lo:lo + self.nbytes].cpu()
        return unpack_entries(part, self.B, self.cap, self.offsets)
