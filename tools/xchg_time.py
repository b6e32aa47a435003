"""Durations of the exchange kernels (push, wait) on the tail stream next to a running head kernel; torchrun, >= 2 ranks."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from yolo_continuous_b200.parallel import PeerExchange
from yolo_continuous_b200.pipeline import PostBackbone

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
head = bench.make_head().to(dev)
pipe = PostBackbone(head, 64, bench.SHAPES, torch.bfloat16, bench.INPUT_SHAPE, bench.IMAGE_SHAPE, True, bench.CONF, bench.IOU, dev,
                    use_graph=False, overlap=True)
xc = PeerExchange(pipe.hdr_ints, 64, 4096, dev)
pipe.attach_exchange(xc)
xs = bench.make_maps(64, 1234 + rank, torch.bfloat16, dev)
for _ in range(5):
    pipe.run_device(xs)
pipe.wait(); xc.wait(pipe.tail_stream); torch.cuda.synchronize(); dist.barrier()
pipe._xchg_events = []
hev = []
for _ in range(40):
    e = tuple(torch.cuda.Event(enable_timing=True) for _ in range(4))
    pipe.run_device(xs, head_events=e)
    hev.append(e)
pipe.wait(); xc.wait(pipe.tail_stream); torch.cuda.synchronize()
ev = pipe._xchg_events[5:]
hev = hev[5:]
if rank == 0:
    print("head ms", statistics.mean(a[0].elapsed_time(a[1]) for a in hev), "nms ms", statistics.mean(a[2].elapsed_time(a[3]) for a in hev))
    print("nms_end->push_start ms", statistics.mean(h[3].elapsed_time(a[0]) for h, a in zip(hev, ev)))
    print("push ms", statistics.mean(a[0].elapsed_time(a[1]) for a in ev), "max", max(a[0].elapsed_time(a[1]) for a in ev))
    print("wait ms", statistics.mean(a[1].elapsed_time(a[2]) for a in ev), "max", max(a[1].elapsed_time(a[2]) for a in ev))
pipe.attach_exchange(None)
xc.close()
dist.destroy_process_group()
