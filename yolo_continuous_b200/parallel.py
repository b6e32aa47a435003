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
