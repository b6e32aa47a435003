"""NMS-only latency at batch 1 (the second metric of BASELINE.json), kernel by kernel (run under ncu for durations)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yolo_continuous_b200 import _lib
from yolo_continuous_b200.pipeline import PostBackbone
dev = torch.device("cuda:0")
head = bench.make_head().to(dev)
g = torch.Generator(device=dev).manual_seed(1234)
xs = [torch.randn(8, c, h, w, generator=g, device=dev).to(torch.bfloat16) for c, (h, w) in zip(bench.CH, bench.SHAPES)]
zsrc = PostBackbone(head, 8, bench.SHAPES, torch.bfloat16, (640, 640), (512, 773), True, 0.25, 0.45, dev, use_graph=False, fused=False)
zsrc.run_device(xs)
for conf, iou in ((0.25, 0.45), (0.001, 0.65)):
    p1 = PostBackbone(head, 1, bench.SHAPES, torch.bfloat16, (640, 640), (512, 773), True, conf, iou, dev, use_graph=False, fused=False)
    p1.z.copy_(zsrc.z[3:4])
    def nms_only():
        m1 = p1.meta.data_ptr()
        _lib.check(_lib.lib.yc_nms_batched(p1.z.data_ptr(), p1.nms_params, p1.ws.data_ptr(), p1.ws.numel(), p1.out_rows.data_ptr(),
                                           p1.out_idx.data_ptr(), m1, m1 + 4, _lib.stream_ptr(dev)), "nms")
    for _ in range(3):
        nms_only()
    torch.cuda.synchronize()
    lat = []
    for i in range(50):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nms_only(); b.record(); b.synchronize()
        lat.append(a.elapsed_time(b))
    print(f"conf {conf}: eager p50 {statistics.median(lat) * 1000:.1f} us, detections {int(p1.meta[0])}")
