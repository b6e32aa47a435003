"""World-size-2 gloo test of the sharding + variable-length detection gather (host-side logic of the
multi-GPU path; the data path itself has no collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolo_continuous_b200.parallel import gather_detections, shard_range
    lo, hi = shard_range(n_images, rank, world)
    g = np.random.default_rng(100 + rank)
    counts = torch.from_numpy(g.integers(0, 5, hi - lo).astype(np.int32))
    if rank == 1:
        counts[:] = 0 if n_images == 4 else counts  # one configuration has an empty rank
    total = int(counts.sum())
    rows = torch.zeros((total + 3, 7))
    rows[:total, 0] = torch.arange(total) + 1000 * rank
    rows[:total, 6] = rank
    got_rows, got_counts = gather_detections(rows, counts)
    ok = len(got_rows) == world
    for r in range(world):
        ok &= bool((got_rows[r][:, 6] == r).all()) and got_rows[r].shape[0] == int(got_counts[r].sum())
        ok &= bool((got_rows[r][:, 0] == torch.arange(got_rows[r].shape[0]) + 1000 * r).all())
    q.put((rank, ok, [int(c.sum()) for c in got_counts]))
    dist.destroy_process_group()


def _run(n_images):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2]      # both ranks agree on everybody's totals


def test_gather_detections_world2():
    _run(10)


def test_gather_detections_with_empty_rank():
    _run(4)


def test_shard_range_covers_everything():
    from yolo_continuous_b200.parallel import shard_range
    for n in (1, 7, 64, 1024):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _worker_fixed(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolo_continuous_b200.parallel import DetectionGather
    bs, cap = 3, 8
    hdr_ints = (2 * bs + 1 + 3) // 4 * 4
    msg = torch.zeros(hdr_ints * 4 + cap * 28, dtype=torch.uint8)
    hdr = msg[:hdr_ints * 4].view(torch.int32)
    counts = torch.tensor([1, 0, 2 + rank], dtype=torch.int32)
    hdr[:bs] = counts
    hdr[bs:2 * bs + 1] = torch.tensor([0] + counts.cumsum(0).tolist(), dtype=torch.int32)
    rows = msg[hdr_ints * 4:].view(torch.float32).view(cap, 7)
    rows[:int(counts.sum()), 0] = torch.arange(int(counts.sum())) + 100.0 * rank
    g = DetectionGather(msg.numel(), "cpu")
    slot = g.gather_async(msg)
    g.wait()
    ok = True
    for r, (c, total, rr) in enumerate(g.unpack(slot, bs, hdr_ints, cap)):
        ok &= c.tolist() == [1, 0, 2 + r] and int(total) == 3 + r
        ok &= rr[:int(total), 0].tolist() == [100.0 * r + i for i in range(3 + r)]
    # grouped exchange: the messages of 3 consecutive steps travel in one collective; a 4th step is flushed alone
    g3 = DetectionGather(msg.numel(), "cpu", every=3)
    slots = []
    for step in range(4):
        rows[0, 1] = 10.0 * step + rank          # something that differs per step
        slots.append(g3.gather_async(msg))
    ok &= slots[:3] == [None, None, 0] and slots[3] is None
    for r, steps in enumerate(g3.unpack(0, bs, hdr_ints, cap)):
        ok &= len(steps) == 3
        for k, (c, total, rr) in enumerate(steps):
            ok &= c.tolist() == [1, 0, 2 + r] and int(total) == 3 + r and float(rr[0, 1]) == 10.0 * k + r
    last = g3.flush()
    ok &= last == 1
    for r, steps in enumerate(g3.unpack(last, bs, hdr_ints, cap, n=1)):
        ok &= float(steps[0][2][0, 1]) == 30.0 + r
    q.put((rank, ok))
    dist.destroy_process_group()


def test_fixed_size_detection_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_fixed, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert all(r[1] for r in res), res
