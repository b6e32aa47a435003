"""Batch sharding and the one collective of the path: a variable-length gather of detections.

Images are independent (the reference's NMS loops per image, detect.py:106), so rank r of W owns a
contiguous block of images and no data-path collective is needed; results are exchanged with one
small all-gather of per-image counts followed by one all-gather of the rows padded to the largest
rank total (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_images, rank, world):
    """Contiguous image block [lo, hi) of `rank`; blocks differ by at most one image."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(rows, counts, group=None):
    """rows [>=total, 7] and counts [bs] of this rank (same device) -> every rank gets
    (list over ranks of rows [total_r, 7], list over ranks of counts [bs_r]).  Ranks may hold different
    numbers of images only if `counts` is padded by the caller to a common length."""
    world = dist.get_world_size(group)
    cnt_all = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(cnt_all, counts, group=group)
    totals = torch.stack([c.sum() for c in cnt_all]).cpu()      # the one host sync of the exchange
    tmax = max(int(totals.max()), 1)
    mine = int(totals[dist.get_rank(group)])
    pad = torch.zeros((tmax, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    pad[:mine] = rows[:mine]
    out = torch.empty((world * tmax, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view(world, tmax, rows.shape[1])
    return [out[r, :int(totals[r])] for r in range(world)], cnt_all


class DetectionGather:
    """Allocation-free, host-sync-free exchange for the steady state: every rank contributes fixed-size messages
    (the header with per-image counts/offsets + the first `gather_rows` detection rows, exactly the prefix
    PostBackbone.message() returns) and one `all_gather_into_tensor` delivers them.  `every` > 1 batches the
    messages of that many consecutive steps into one collective: a NCCL kernel cannot share an SM with a resident
    head CTA and takes 20-130 us next to a running head kernel, so one collective per step would set the pace of the
    pipeline (measured on 2 B200: 132 us per step against 89 us for the kernels), one per 8 steps does not.
    On CUDA the collective runs on a side stream (or on `stream`, e.g. PostBackbone.tail_stream, in order behind the
    kernels that produced the message), so the exchange of one group of steps overlaps the next steps.  A rank
    whose detections exceed `gather_rows` is visible in its header (offsets[-1] > gather_rows); callers fetch the
    remainder with `gather_detections`."""

    def __init__(self, msg_bytes, device, group=None, n_bufs=2, every=1):
        self.group, self.world = group, dist.get_world_size(group)
        self.msg_bytes, self.every = msg_bytes, int(every)
        self.stage = [torch.empty((self.every * msg_bytes,), dtype=torch.uint8, device=device) for _ in range(n_bufs)]
        self.out = [torch.empty((self.world * self.every * msg_bytes,), dtype=torch.uint8, device=device)
                    for _ in range(n_bufs)]
        self.slot, self.fill = 0, 0
        self.cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.copy_stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.copy_done = None
        self.last_stream = None
        self.done = [None] * n_bufs

    def gather_async(self, msg, stream=None):
        """Queue `msg`; after `every` calls the group is exchanged.  Returns the slot holding the group's result once
        the exchange was issued, else None.  stream: the stream that produced `msg` (e.g. PostBackbone.tail_stream;
        default: the current stream).  The copy into the group buffer runs in order on that stream; the collective
        runs on this object's own stream behind an event, so it never delays the producer's next kernels."""
        assert msg.numel() == self.msg_bytes and msg.dtype == torch.uint8
        part = self.stage[self.slot][self.fill * self.msg_bytes:(self.fill + 1) * self.msg_bytes]
        first = self.fill == 0
        self.fill += 1
        send = self.fill == self.every
        if not self.cuda:
            part.copy_(msg)
            return self._send() if send else None
        from . import _lib
        if stream is None:
            # `msg` was produced on the current stream, which should go straight on to the next step: the staging copy
            # runs on a copy stream behind an event, and the current stream only waits for the copy of the PREVIOUS
            # call (long finished) so that a message buffer is never overwritten while its copy is pending
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            stream = self.copy_stream
            stream.wait_event(ev)
            wait_prev = self.copy_done
        else:
            main, wait_prev = None, None
        self.last_stream = stream
        if first and self.done[self.slot] is not None:
            stream.wait_event(self.done[self.slot])    # the collective that last read this group buffer
        # one driver call: a framework copy under a stream context costs ~15 us of host time per step
        _lib.check(_lib.lib.yc_copy_async(part.data_ptr(), msg.data_ptr(), self.msg_bytes, stream.cuda_stream), "yc_copy_async")
        if main is not None:
            self.copy_done = torch.cuda.Event()
            self.copy_done.record(stream)
            if wait_prev is not None:
                main.wait_event(wait_prev)
        return self._send_cuda(stream) if send else None

    def _send_cuda(self, producer):
        ev = torch.cuda.Event()
        ev.record(producer)
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            slot = self._send()
            self.done[slot] = torch.cuda.Event()
            self.done[slot].record(self.stream)
        return slot

    def _send(self):
        n = self.fill
        slot = self.slot
        src = self.stage[slot][:n * self.msg_bytes]
        dst = self.out[slot][:self.world * n * self.msg_bytes]
        dist.all_gather_into_tensor(dst, src, group=self.group)
        self.slot, self.fill = (slot + 1) % len(self.out), 0
        return slot

    def flush(self, stream=None):
        """Exchange a partially filled group (every rank must call it at the same point)."""
        if self.fill == 0:
            return None
        if not self.cuda:
            return self._send()
        return self._send_cuda(stream if stream is not None else self.last_stream)

    def wait(self):
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)

    def unpack(self, slot, bs, hdr_ints, gather_rows, n=None):
        """-> list over ranks of lists over the group's steps of (counts [bs], total, rows [gather_rows, 7]) views
        (no copy); with every == 1 the inner list is dropped (one tuple per rank)."""
        n = self.every if n is None else n
        res = []
        for r in range(self.world):
            steps = []
            for k in range(n):
                o = (r * n + k) * self.msg_bytes
                m = self.out[slot][o:o + self.msg_bytes]
                hdr = m[:hdr_ints * 4].view(torch.int32)
                rows = m[hdr_ints * 4:].view(torch.float32).view(gather_rows, 7)
                steps.append((hdr[:bs], hdr[2 * bs], rows))
            res.append(steps[0] if self.every == 1 else steps)
        return res


class _DevMem:
    """A device allocation made by the C library, viewed by torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerExchange:
    """The detection exchange over NVLink peer memory (include/yc_b200.h, yc_xchg_*): per step ONE kernel of this rank
    stores its message (header + the first `gather_rows` detection rows) into a slot of every rank's receive buffer and
    raises a flag there; a second one-warp kernel waits for the flags of all ranks.  No NCCL call, no host work and no host
    synchronisation per step; sequence numbers live on the device, so both kernels are captured inside the per-step CUDA
    graph of PostBackbone (attach with `PostBackbone.attach_exchange`).  torch.distributed is used once, to hand the
    CUDA IPC handles around.  A rank whose detections exceed `gather_rows` says so in its header
    (offsets[-1] > gather_rows); callers fetch the remainder with `gather_detections`.

    wait() for sequence number j declares every message before j consumed (their slots may be overwritten): read the
    views of `unpack()` before waiting `slots - 1` steps further."""

    def __init__(self, hdr_ints, bs, gather_rows, device, group=None, slots=8):
        import ctypes as C

        from . import _lib
        self._lib, self._C = _lib, C
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.hdr_ints, self.bs, self.gather_rows, self.slots = int(hdr_ints), int(bs), int(gather_rows), int(slots)
        self.msg_bytes = (self.hdr_ints * 4 + self.gather_rows * 28 + 15) // 16 * 16
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            buf = C.c_void_p()
            handle = (C.c_uint8 * 64)()
            _lib.check(_lib.lib.yc_xchg_alloc(self.world, self.slots, self.msg_bytes, C.byref(buf), handle), "yc_xchg_alloc")
            self._buf = buf.value
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self._opened, ptrs = [], []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(self._buf)
                    continue
                p = C.c_void_p()
                hb = (C.c_uint8 * 64).from_buffer_copy(h)
                _lib.check(_lib.lib.yc_xchg_open(hb, C.byref(p)), f"yc_xchg_open (rank {r}: CUDA IPC / peer access)")
                self._opened.append(p.value)
                ptrs.append(p.value)
            self.peers = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
            nbytes = _lib.lib.yc_xchg_bytes(self.world, self.slots, self.msg_bytes)
            self._mem = _DevMem(self._buf, nbytes)
            self.recv = torch.as_tensor(self._mem, device=self.device)   # uint8 view of this rank's receive buffer
            dist.barrier(group=group)    # every rank has opened every buffer before anybody pushes

    def push(self, msg, stream=None):
        """Enqueue the push of `msg` (uint8 tensor: the prefix PostBackbone.message() returns, or the whole buffer).  The
        i-th push of a rank (counting replays of a graph that captured it) carries sequence number i."""
        s = self._C.c_void_p((stream or torch.cuda.current_stream(self.device)).cuda_stream)
        self._lib.check(self._lib.lib.yc_xchg_push(msg.data_ptr(), self.hdr_ints, self.bs, self.gather_rows, self.peers.data_ptr(),
                                                   self.world, self.rank, self.slots, self.msg_bytes, s), "yc_xchg_push")

    def wait(self, stream=None, lag=0):
        """Enqueue the wait for this rank's next un-awaited sequence number j: work queued behind it on `stream` sees the
        messages of all ranks with that number.  lag = 1: a no-op unless this rank has pushed message j + 1 already
        ("push(i); wait(lag=1)" every step awaits step i - 1 and never spins in the steady state)."""
        s = self._C.c_void_p((stream or torch.cuda.current_stream(self.device)).cuda_stream)
        self._lib.check(self._lib.lib.yc_xchg_wait(self.peers.data_ptr(), self.world, self.rank, self.slots, self.msg_bytes,
                                                   int(lag), s), "yc_xchg_wait")

    def wait_all(self, stream=None):
        """Enqueue the wait for every message this rank has pushed so far (end of a stream of batches)."""
        self.wait(stream, lag=-1)

    def state(self):
        """(next push seq, next wait seq, error) read back from the device (synchronises the current stream)."""
        out = (self._C.c_uint32 * 4)()
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib.yc_xchg_state(self._buf, self.world, self.slots, self.msg_bytes, out,
                                                        self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                            "yc_xchg_state")
        return int(out[0]), int(out[1]), int(out[3])

    def unpack(self, seq):
        """-> list over ranks of (counts [bs], total, rows [gather_rows, 7]) views of this rank's receive buffer for the
        messages with sequence number `seq` (no copy; valid until wait() has been called slots - 1 more times)."""
        slot = seq % self.slots
        res = []
        for r in range(self.world):
            o = (slot * self.world + r) * self.msg_bytes
            m = self.recv[o:o + self.msg_bytes]
            hdr = m[:self.hdr_ints * 4].view(torch.int32)
            rows = m[self.hdr_ints * 4:self.hdr_ints * 4 + self.gather_rows * 28].view(torch.float32).view(self.gather_rows, 7)
            res.append((hdr[:self.bs], hdr[2 * self.bs], rows))
        return res

    def close(self):
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            for p in self._opened:
                self._lib.lib.yc_xchg_close(self._C.c_void_p(p))
            self._opened = []
            dist.barrier(group=self.group)
            if self._buf:
                self.recv = None
                self._lib.lib.yc_xchg_free(self._C.c_void_p(self._buf))
                self._buf = None
