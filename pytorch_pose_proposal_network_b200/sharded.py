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


class PeerPoseGatherer:
    """Gather of every rank's poses at ONE root rank over peer memory — no collective on the data plane.

    ``north_star``: whole images are sharded over the GPUs of one box, "NCCL is used only to gather per-rank pose
    lists and timings".  The pose lists are small (≈ 1 MB per 512 images) but a collective per step — or one large
    one per group of steps — costs SM time under the arg-max stream and an exposed tail at the end of a run.  Here
    the root rank owns one buffer ``[world][slots][packed record]`` allocated by the library (``ppn_peer_alloc``);
    its CUDA IPC handle goes round once through the process group and every rank maps it (``ppn_peer_open``:
    NVLink peer access).  Then, per step,

    * ``mode="store"`` (default): the parse kernel writes its dense (human, part) records straight into this
      rank's slot of the ROOT's buffer (``ppn_parse_dense_remote``): the compute step and the gather are one
      kernel, the records cross NVLink as plain stores while the kernel runs, exactly as many bytes as were
      produced;
    * ``mode="copy"``: the records go to a local slot and every ``notify_every`` steps ONE ``cudaMemcpyAsync``
      (copy engine, no SM) on a side stream ships the group's slots to the root.

    The control plane — "have the records landed?" — has two forms (``control``):

    * ``"flags"`` (default without a consumer): no collective at all.  Behind the steps it covers, every rank stores
      its step counter into ITS 64-bit slot of a counter array at the end of the root's buffer (``ppn_peer_post``: a
      one-thread kernel on the side stream, fully ordered behind the parse kernels, release at system scope); at
      ``finish()`` every rank posts once more with the run number in the counter's upper half, and the root's stream
      waits until all counters carry it (``ppn_peer_wait``: one polling warp, with a timeout) — however many steps each
      rank parsed.  A rank reuses a slot ``slots`` steps later — long after its own kernel that
      wrote it has completed (stream order), so nothing else is needed while nobody consumes the records between
      steps.  At the driver's 20 steps the exposed tail of a run is one store per rank instead of an NCCL collective;
    * ``"nccl"`` (round 2's first form; required with a ``consumer``, whose progress gates the reuse of slots):

    every ``notify_every`` steps an 8-byte ``all_gather`` of step
    counters on the side stream, ordered after the steps it covers.  When it completes on the root, every rank's
    records of those steps have landed (a kernel's stores are visible when it has completed; the copy is on the
    same stream); a consumer's work enqueued on the root's side stream right after it therefore reads complete
    data, and — being stream-ordered before the root's NEXT notification — is finished before any rank, having
    seen that next notification complete, reuses a slot.  Hence ``slots >= 2 * notify_every``.
    """

    def __init__(self, parser, images_per_rank: int, cap_entries: int, group=None, slots: int = 32,
                 notify_every: int = 8, root: int = 0, mode: str = "store", consumer=None, control: str = None,
                 timeout_ms: int = 5000):
        import ctypes as C
        from . import _lib
        if mode not in ("store", "copy"):
            raise ValueError("mode must be 'store' or 'copy'")
        control = control or ("nccl" if consumer is not None else "flags")
        if control not in ("flags", "nccl"):
            raise ValueError("control must be 'flags' or 'nccl'")
        if control == "flags" and consumer is not None:
            raise ValueError("a consumer needs control='nccl': its progress gates the reuse of slots")
        self.control, self.timeout_ms = control, int(timeout_ms)
        self.parser, self.group, self.mode, self.root = parser, group, mode, int(root)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.B, self.cap = int(images_per_rank), int(cap_entries)
        self.ne = max(1, int(notify_every))
        self.slots = max(int(slots), 2 * self.ne)
        self.consumer = consumer
        self.nbytes, self.offsets = parser.packed_layout(self.B, self.cap)
        self.lib = parser.lib
        dev = parser.device
        region = self.slots * self.nbytes
        self._owned = None
        handle = None
        with torch.cuda.device(dev):
            if self.rank == self.root:
                ptr, buf = C.c_void_p(), C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
                _lib.check(self.lib.ppn_peer_alloc(self.world * region + 8 * self.world + 256, C.byref(ptr), buf), "ppn_peer_alloc")
                self._owned, handle = ptr.value, buf.raw
            handles = [None] * self.world
            dist.all_gather_object(handles, handle, group=group)
            if self.rank == self.root:
                self.base = self._owned
            else:
                ptr = C.c_void_p()
                _lib.check(self.lib.ppn_peer_open(handles[self.root], C.byref(ptr)), "ppn_peer_open")
                self.base = ptr.value
        self.region = region
        self.mine_at = self.base + self.rank * region
        self.flags_at = self.base + -(-self.world * region // 256) * 256     # int64 [world]: the landing counters (root's memory)
        self.timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
        if control == "flags":
            if self.rank == self.root:                                       # cleared before anybody can post
                zeros = torch.zeros(self.world, dtype=torch.int64, device=dev)
                with torch.cuda.device(dev):
                    _lib.check(self.lib.ppn_peer_copy(self.flags_at, zeros.data_ptr(), 8 * self.world,
                                                      torch.cuda.current_stream(dev).cuda_stream), "ppn_peer_copy")
                torch.cuda.synchronize(dev)
            dist.barrier(group=group)
        hdr = -(-4 * (2 + 3 * self.B) // 256) * 256
        self.local_stride = self.nbytes if mode == "copy" else hdr
        self.local = torch.zeros(self.slots * self.local_stride, dtype=torch.uint8, device=dev)
        self._local_slices = [self.local[k * self.local_stride:(k + 1) * self.local_stride] for k in range(self.slots)]
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.seen = torch.zeros(self.world, dtype=torch.int64, device=dev)
        self.side = torch.cuda.Stream(device=dev)
        self._events = [torch.cuda.Event() for _ in range(4)]
        self._notes = []                      # notification n -> work handle (None once waited for)
        self.step = 0
        self._covered = 0                     # steps [0, _covered) are covered by an issued notification
        self._epoch = 0                       # finish() calls so far (flags control: upper half of the landing counters)

    # ---- control plane ------------------------------------------------------------------------------
    def _notify(self):
        """Ship (copy mode) and announce every step enqueued so far."""
        upto, dev = self.step, self.parser.device
        main = torch.cuda.current_stream(dev)
        ev = self._events[len(self._notes) % len(self._events)]
        ev.record(main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            if self._notes and self._notes[-1] is not None:
                self._notes[-1].wait()         # side stream: the previous notification has read `counter` / written `seen`
            if self.mode == "copy":
                s = self._covered
                while s < upto:                # contiguous runs of slots (a run ends where the ring wraps)
                    k = s % self.slots
                    n = min(upto - s, self.slots - k)
                    with torch.cuda.device(dev):
                        rc = self.lib.ppn_peer_copy(self.mine_at + k * self.nbytes, self.local.data_ptr() + k * self.nbytes,
                                                    n * self.nbytes, self.side.cuda_stream)
                    if rc:
                        from . import _lib
                        raise _lib.PPNError(rc, "ppn_peer_copy")
                    s += n
            if self.control == "flags":
                with torch.cuda.device(dev):
                    rc = self.lib.ppn_peer_post(self.flags_at + 8 * self.rank, (self._epoch << 32) | upto, self.side.cuda_stream)
                if rc:
                    from . import _lib
                    raise _lib.PPNError(rc, "ppn_peer_post")
                if self.mode == "copy":        # the local slots of these steps may be rewritten once this copy has run
                    self._notes.append(_Done(self.side))
                self._covered = upto
                return
            self.counter.fill_(upto)
            work = dist.all_gather_into_tensor(self.seen, self.counter, group=self.group, async_op=True)
            if self.consumer is not None and self.rank == self.root:
                work.wait()                    # side stream: the consumer's kernels run after the records have landed
                self.consumer(self, self._covered, upto)
                work = _Done(self.side)
        self._notes.append(work)
        self._covered = upto

    def _wait_note(self, n: int):
        """Make the CURRENT stream wait for notification n (and all earlier ones)."""
        for i in range(min(n, len(self._notes) - 1) + 1):
            w = self._notes[i]
            if w is not None:
                w.wait()
                self._notes[i] = None

    # ---- the step -----------------------------------------------------------------------------------
    def parse(self, head, out=None, input_complete: bool = False):
        s = self.step
        k = s % self.slots
        if s >= self.slots and (self.control == "nccl" or self.mode == "copy"):
            # the slot still holds step s - slots: see the class comment (flags + store: nothing to wait for — the kernel
            # that wrote it has long completed; flags + copy: the side-stream copy that shipped it must have run)
            self._wait_note((s - self.slots) // self.ne + 1)
        if self.mode == "store":
            res = self.parser.parse(head, out=out, input_complete=input_complete, dense=self._local_slices[k], cap_entries=self.cap,
                                    skip_slots=True, remote=(self.mine_at + k * self.nbytes, self.nbytes))
        else:
            res = self.parser.parse(head, out=out, input_complete=input_complete, dense=self._local_slices[k], cap_entries=self.cap,
                                    skip_slots=True)
        self.step += 1
        if self.step % self.ne == 0:
            self._notify()
        return res

    def finish(self):
        """Announce the remaining steps and make the current stream wait until every rank's records have landed."""
        if self.control == "flags":
            # the counters carry (number of finish() calls << 32 | steps): the root waits for every rank to have finished
            # THIS run, however many steps each of them parsed (uneven shards)
            self._epoch += 1
            self._notify()
            dev = self.parser.device
            if self.rank == self.root:
                with torch.cuda.stream(self.side), torch.cuda.device(dev):
                    rc = self.lib.ppn_peer_wait(self.flags_at, self.world, self._epoch << 32, self.timeout_ms,
                                                self.timed_out.data_ptr(), self.side.cuda_stream)
                if rc:
                    from . import _lib
                    raise _lib.PPNError(rc, "ppn_peer_wait")
            torch.cuda.current_stream(dev).wait_stream(self.side)
            return
        if self.step > self._covered:
            self._notify()
        self._wait_note(len(self._notes) - 1)

    def check_landed(self):
        """After finish() and a synchronize: raise if the root gave up waiting for a rank's landing counter."""
        if int(self.timed_out.item()):
            raise RuntimeError(f"peer gather: a rank's records had not landed after {self.timeout_ms} ms")

    # ---- reading (root) -----------------------------------------------------------------------------
    def records_of(self, rank: int, step_back: int = 0):
        """Host view of `rank`'s records of the last parsed step minus `step_back` (root only; after finish())."""
        from .parser import unpack_entries
        if self.rank != self.root:
            raise RuntimeError("the records are gathered at the root rank only")
        s = self.step - 1 - step_back
        held = self.slots if self.control == "flags" else self.slots - 2 * self.ne + 1     # flags: a slot is only rewritten `slots` steps later
        if s < 0 or step_back < 0 or step_back >= held:
            raise ValueError("that step's slot may already have been reused")
        at = rank * self.region + (s % self.slots) * self.nbytes
        from . import _lib
        host = torch.empty(self.nbytes, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize(self.parser.device)
        with torch.cuda.device(self.parser.device):
            _lib.check(self.lib.ppn_peer_copy(host.data_ptr(), self.base + at, self.nbytes, None), "ppn_peer_copy")
        torch.cuda.synchronize(self.parser.device)
        return unpack_entries(host, self.B, self.cap, self.offsets, derive=(self.mode == "store"))

    def close(self):
        torch.cuda.synchronize(self.parser.device)
        with torch.cuda.device(self.parser.device):
            if self._owned:
                dist.barrier(group=self.group)            # peers unmap first
                self.lib.ppn_peer_free(self._owned)
                self._owned = None
            elif self.base:
                self.lib.ppn_peer_close(self.base)
                dist.barrier(group=self.group)
            self.base = 0


class _Done:
    """Stands in for a finished collective: waiting makes the current stream wait for `stream` as it is now."""

    def __init__(self, stream):
        self.ev = torch.cuda.Event()
        self.ev.record(stream)

    def wait(self):
        torch.cuda.current_stream().wait_event(self.ev)
