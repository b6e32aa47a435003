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
