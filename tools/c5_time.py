"""C5 fused step (head -> NMS candidates) of IAuxDetect / IBin at 1280x1280, 16 images: head-kernel time through
PostBackbone.run_device with events around the head call; used with the YC_TC_DEBUG switches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yolo_continuous_b200.nets import IAuxDetect, IBin
from yolo_continuous_b200.pipeline import PostBackbone

dev = torch.device("cuda:0")
shapes5 = [(160, 160), (80, 80), (40, 40)]
for name in os.environ.get("CASES", "ibin,iaux").split(","):
    cls, ch = (IAuxDetect, bench.CH * 2) if name == "iaux" else (IBin, bench.CH)
    head = cls(bench.NC, bench.COCO_ANCHORS, ch).eval()
    with torch.no_grad():
        for im in head.im:
            im.implicit.copy_(1.0 + 0.02 * torch.randn(im.implicit.shape, generator=torch.Generator().manual_seed(11)))
        g5 = torch.Generator().manual_seed(12)
        for conv in head.m:
            k = conv.weight.shape[1]
            w = conv.weight.view(head.na, head.no, k)
            w[:, head.no - bench.NC - 1:, :] = torch.randn(head.na, bench.NC + 1, k, generator=g5) * (1.5 / k ** 0.5)
            b = conv.bias.view(head.na, head.no)
            b[:, head.no - bench.NC - 1] = -5.0
            b[:, head.no - bench.NC:] = -3.0
    head = head.to(dev)
    head.stride = torch.tensor(bench.STRIDES)
    xs = bench.make_maps(16, 7, torch.bfloat16, dev, ch, shapes5 * (len(ch) // 3))[:3]
    pipe = PostBackbone(head, 16, shapes5, torch.bfloat16, (1280, 1280), (720, 1280), True, bench.CONF, bench.IOU, dev,
                        use_graph=False, overlap=False)
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(8)]
    for e in ev:
        pipe.run_device(xs, head_events=e)
    pipe.wait()
    torch.cuda.synchronize()
    hm = sorted(e[0].elapsed_time(e[1]) for e in ev)[len(ev) // 2]
    nm = sorted(e[2].elapsed_time(e[3]) for e in ev)[len(ev) // 2]
    print(f"TC={os.environ.get('YC_TC_DEBUG', '0')} {name}: fused head kernel {hm * 1e3:.1f} us, NMS kernels {nm * 1e3:.1f} us, "
          f"detections {int(pipe.meta[16:][-1]) if hasattr(pipe, 'meta') else -1}", flush=True)
    del pipe
