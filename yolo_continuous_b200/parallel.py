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
    """Allocation-free, host-sync-free exchange for the steady state: every rank contributes one fixed-size
    message (the header with per-image counts/offsets + the first `gather_rows` detection rows, exactly the
    prefix PostBackbone.message() returns) and one `all_gather_into_tensor` delivers all of them.  On CUDA
    the collective runs on a side stream behind an event, so with double-buffered outputs the exchange of
    step i overlaps step i+1.  A rank whose detections exceed `gather_rows` is visible in its header
    (offsets[-1] > gather_rows); callers fetch the remainder with `gather_detections`."""

    def __init__(self, msg_bytes, device, group=None, n_bufs=2):
        self.group, self.world = group, dist.get_world_size(group)
        self.msg_bytes = msg_bytes
        self.out = [torch.empty((self.world * msg_bytes,), dtype=torch.uint8, device=device) for _ in range(n_bufs)]
        self.slot = 0
        self.cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None

    def gather_async(self, msg, stream=None):
        """stream: run the collective in order on this stream (e.g. PostBackbone.tail_stream, which produced
        `msg`) instead of this object's own side stream behind an event on the current stream."""
        assert msg.numel() == self.msg_bytes and msg.dtype == torch.uint8
        self.slot = (self.slot + 1) % len(self.out)
        dst = self.out[self.slot]
        if self.cuda and stream is not None:
            self.last_stream = stream
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(dst, msg, group=self.group)
        elif self.cuda:
            self.last_stream = self.stream
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                dist.all_gather_into_tensor(dst, msg, group=self.group)
        else:
            dist.all_gather_into_tensor(dst, msg, group=self.group)
        return self.slot

    def wait(self):
        if self.cuda:
            torch.cuda.current_stream().wait_stream(getattr(self, "last_stream", self.stream))

    def unpack(self, slot, bs, hdr_ints, gather_rows):
        """-> list over ranks of (counts [bs], total, rows [min(total, gather_rows), 7]) views (no copy)."""
        res = []
        for r in range(self.world):
            m = self.out[slot][r * self.msg_bytes:(r + 1) * self.msg_bytes]
            hdr = m[:hdr_ints * 4].view(torch.int32)
            rows = m[hdr_ints * 4:].view(torch.float32).view(gather_rows, 7)
            res.append((hdr[:bs], hdr[2 * bs], rows))
        return res
