"""Timing of the other BASELINE.json configurations on one B200 (they are parity-test cases, not bench lines):
C3  mAP-eval stress: conf 0.001 / iou 0.65, batch 256, reference-init ImplicitM (every row a candidate) and trained-like
C5  IAuxDetect and IBin at 1280x1280 (100 800 rows per image), 16 images per GPU
"""
import json, sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yolo_continuous_b200.nets import IAuxDetect, IBin
from yolo_continuous_b200.pipeline import PostBackbone

dev = torch.device("cuda:0")


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = {}
# ---- C3 ----------------------------------------------------------------------------------------------------
for name, ref_init in (("c3_trained_like_conf0.001", False), ("c3_stress_dense_ref_init_im", True)):
    head = bench.make_head().to(dev)
    if ref_init:
        g = torch.Generator().manual_seed(3)
        with torch.no_grad():
            for m in head.im:
                m.implicit.copy_((torch.randn(m.implicit.shape, generator=g) * 0.02).to(dev))   # nets/common.py:430
    bs = 256
    g = torch.Generator(device=dev).manual_seed(1234)
    xs = [torch.randn(bs, c, h, w, generator=g, device=dev).to(torch.bfloat16) for c, (h, w) in zip(bench.CH, bench.SHAPES)]
    pipe = PostBackbone(head, bs, bench.SHAPES, torch.bfloat16, (640, 640), (512, 773), True, 0.001, 0.65, dev,
                        use_graph=False, overlap=True)
    ms = timed(lambda: pipe.submit(xs), n=10, warm=2)
    pipe.drain()
    torch.cuda.synchronize()
    out[name] = {"bs": bs, "ms_per_step": ms, "images_per_s": bs / ms * 1e3, "detections": int(pipe.meta[bs:][-1]),
                 "candidates_per_image": None}
    del pipe, xs
    torch.cuda.empty_cache()

# ---- C5 ----------------------------------------------------------------------------------------------------
shapes = [(160, 160), (80, 80), (40, 40)]
bs = 16
g = torch.Generator(device=dev).manual_seed(7)
for name, cls, ch in (("c5_iaux_1280_bf16", IAuxDetect, bench.CH * 2), ("c5_ibin_1280_bf16", IBin, bench.CH)):
    head = cls(80, bench.COCO_ANCHORS, ch).to(dev).eval()
    head.stride = torch.tensor(bench.STRIDES)
    head.return_raw = False
    head.compute_aux_in_eval = False   # the reference's dead aux convolution in eval (nets/iaux_detect.py:37-38)
    xs = [torch.randn(bs, c, h, w, generator=g, device=dev).to(torch.bfloat16) for c, (h, w) in zip(ch, shapes * (len(ch) // 3))]
    with torch.no_grad():
        ms = timed(lambda: head(list(xs)), n=10, warm=2)
        z = head(list(xs))[0]
    rows = z.shape[1]
    flops = 2 * head.na * head.no * sum(c * h * w for c, (h, w) in zip(ch[:3], shapes)) * bs
    bytes_ = (sum(c * h * w for c, (h, w) in zip(ch[:3], shapes)) * 2 + rows * z.shape[2] * 4) * bs
    out[name] = {"bs": bs, "rows_per_image": rows, "forward_ms": ms, "images_per_s": bs / ms * 1e3,
                 "tflops": flops / ms / 1e9, "hbm_gbs_algorithmic": bytes_ / ms / 1e6}
print(json.dumps(out, indent=1))

# ---- C1: yolov7-tiny head (nc = 1, ch 128/256/512), one 640x640 image, conf / iou 0.3 (detect.py:271-272) ---------
import time

from yolo_continuous_b200 import detect as b200
from yolo_continuous_b200.nets import IDetect
from oracle import ref_port
TINY_ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
ch1 = (128, 256, 512)
g1 = torch.Generator().manual_seed(0)
h1 = IDetect(1, TINY_ANCHORS, ch1).eval()
with torch.no_grad():
    for n_, p_ in h1.named_parameters():
        if n_.endswith("weight"):
            p_.copy_(0.02 * torch.randn(p_.shape, generator=g1))
        elif n_.startswith("im."):
            p_.copy_(1.0 + 0.02 * torch.randn(p_.shape, generator=g1))
        elif n_.endswith("bias"):
            p_.copy_(torch.full(p_.shape, -2.0))
h1.stride = torch.tensor(bench.STRIDES)
xs_cpu = [torch.randn(1, c, h, w, generator=g1) for c, (h, w) in zip(ch1, bench.SHAPES)]
params = {"anchors": h1.anchor_grid.detach().reshape(3, -1, 2).numpy(), "w": [m.weight.detach()[:, :, 0, 0] for m in h1.m],
          "b": [m.bias.detach() for m in h1.m], "ia": [a.implicit.detach().reshape(-1) for a in h1.ia],
          "im": [m.implicit.detach().reshape(-1) for m in h1.im]}
ts = []
with torch.no_grad():
    for i in range(7):
        t0 = time.perf_counter()
        ref = ref_port.post_backbone(params, [x.clone() for x in xs_cpu], bench.STRIDES, 1, (640, 640), (512, 773), True, 0.3, 0.3)
        ts.append(time.perf_counter() - t0)
cpu_ms = statistics.median(ts[2:]) * 1e3
h1 = h1.to(dev)
out_c1 = {}
for dt in (torch.float32, torch.bfloat16):
    xs1 = [x.to(dev).to(dt) for x in xs_cpu]
    lat = []
    with torch.no_grad():
        for i in range(30):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = b200.detect_post_backbone(h1, list(xs1), (640, 640), (512, 773), True, 0.3, 0.3)
            lat.append(time.perf_counter() - t0)
    out_c1[str(dt)] = {"post_backbone_ms_p50_host_clock": statistics.median(lat[5:]) * 1e3,
                       "detections": 0 if res[0] is None else len(res[0])}
out_c1["reference_cpu_port_ms"] = cpu_ms
out_c1["reference_cpu_detections"] = 0 if ref[0] is None else len(ref[0])
print(json.dumps({"c1_tiny_bs1": out_c1}, indent=1))
