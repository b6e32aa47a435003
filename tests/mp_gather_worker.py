"""Two-rank worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU, NCCL):
every rank runs the fused step on ITS images and the detections are exchanged (a) with the fixed-size grouped
all-gather of the steady state (DetectionGather), (b) with the variable-length gather_detections that serves ranks whose
detections exceed the fixed message (overflow path).  Every rank recomputes every other rank's local result from the
seeded inputs and checks that what arrived is bit-identical to what that rank produced."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from test_gpu_parity import _bench_like_head
    from yolo_continuous_b200.parallel import DetectionGather, gather_detections
    from yolo_continuous_b200.pipeline import PostBackbone
    ch, shapes, bs = (64, 128, 256), [(40, 40), (20, 20), (12, 12)], 4
    head = _bench_like_head(80, ch, 3).to(dev)
    pipe = PostBackbone(head, bs, shapes, torch.bfloat16, (320, 320), (240, 320), True, 0.25, 0.45, dev, use_graph=False,
                        overlap=True)
    n_steps = 5

    def inputs(r, step):
        g = torch.Generator(device=dev).manual_seed(1000 * r + step)
        return [torch.randn(bs, c, h, w, generator=g, device=dev).to(torch.bfloat16) for c, (h, w) in zip(ch, shapes)]

    # what every rank produces at every step (recomputed locally: the step is deterministic)
    local_res = {}
    for r in range(world):
        for s in range(n_steps):
            rows, idx, counts, offsets = pipe.run_device(inputs(r, s))
            pipe.wait()
            torch.cuda.synchronize()
            tot = int(offsets[-1])
            local_res[(r, s)] = (counts.clone(), rows[:tot].clone())
    assert min(int(v[0].sum()) for v in local_res.values()) > 8, "test needs detections on every rank"

    for gather_rows in (4096, 8):      # 8 rows: every rank overflows the fixed message
        gat = DetectionGather(pipe.message(gather_rows).numel(), dev, every=2)
        slots = []
        for s in range(n_steps):       # pipelined graph form, as bench.py runs it
            prev = pipe.submit(inputs(rank, s))
            if prev is not None:
                slots.append(gat.gather_async(pipe.message(gather_rows, previous=True)))
        pipe.drain()
        slots.append(gat.gather_async(pipe.message(gather_rows)))
        slots.append(gat.flush())
        gat.wait()
        torch.cuda.synchronize()
        # groups of 2 steps: slots[1] -> steps 0,1 ; slots[3] -> steps 2,3 ; flush -> step 4
        assert [x is not None for x in slots] == [False, True, False, True, False, True], slots
        # the group buffers alternate (2 of them): steps 2,3 and the flushed step 4 are still there
        for slot, steps, n in ((slots[3], (2, 3), None), (slots[5], (4,), 1)):
            for r, got in enumerate(gat.unpack(slot, bs, pipe.hdr_ints, gather_rows, n=n)):
                for k, s in enumerate(steps):
                    counts, total, rows = got[k]
                    want_c, want_r = local_res[(r, s)]
                    assert torch.equal(counts, want_c), (gather_rows, r, s)
                    assert int(total) == want_r.shape[0]
                    m = min(int(total), gather_rows)
                    assert torch.equal(rows[:m], want_r[:m]), (gather_rows, r, s)
                    if int(total) > gather_rows and s == 4:
                        # overflow: the header says so; the remainder comes with the variable-length gather
                        rows_l, _, counts_l, _ = pipe._views(pipe.cur)
                        all_rows, all_counts = gather_detections(rows_l, counts_l)
                        for q in range(world):
                            assert torch.equal(all_counts[q], local_res[(q, 4)][0])
                            assert torch.equal(all_rows[q], local_res[(q, 4)][1])
    # ---- the repo's own exchange: push / wait kernels over NVLink peer memory, inside the per-step graphs ------------
    from yolo_continuous_b200.parallel import PeerExchange
    for gather_rows in (4096, 8):
        xc = PeerExchange(pipe.hdr_ints, bs, gather_rows, dev, slots=4 if gather_rows == 8 else 8)
        pipe.attach_exchange(xc)
        n_rounds = 3                   # 15 messages: every slot is reused three times (credits / acknowledgements)
        for rnd in range(n_rounds):
            for s in range(n_steps):
                pipe.submit(inputs(rank, s))
                if rank == 1 and s == 2:
                    torch.cuda._sleep(20_000_000)    # one rank falls behind: the others must wait, not overwrite
        pipe.drain()
        xc.wait_all()
        sp, sw, err = xc.state()
        assert (sp, sw, err) == (n_rounds * n_steps, n_rounds * n_steps, 0), (sp, sw, err)
        # the last `slots` messages are still in the receive buffer: sequence number q holds step q % n_steps
        for q in range(sp - 3, sp):
            for r, (counts, total, rows) in enumerate(xc.unpack(q)):
                want_c, want_r = local_res[(r, q % n_steps)]
                assert torch.equal(counts, want_c), ("peer", gather_rows, r, q)
                assert int(total) == want_r.shape[0]
                m = min(int(total), gather_rows)
                assert torch.equal(rows[:m], want_r[:m]), ("peer", gather_rows, r, q)
        pipe.attach_exchange(None)
        xc.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_GATHER_OK PEER_EXCHANGE_OK world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
