#!/usr/bin/env python
"""bench.py -- post-backbone images/sec of the B200 path (and the reference's CPU path beside it).

Workload (BASELINE.json configs[1], "C2"): yolov7 COCO IDetect head (80 classes, 3 anchors x 3 strides,
ch 256/512/1024) at 640x640, batch 64 per GPU, synthetic feature maps, seeded trained-like weights,
conf 0.25 / iou 0.45.  One "step" = one pass of the hot path over one batch:
    head 1x1 conv (+ImplicitA/M) -> sigmoid + box decode -> conf threshold + compaction -> per-class NMS
    -> letterbox undo.
`value` is measured with inputs resident in HBM over exactly K steps; `sustained` repeats the same step for >= 2 s
(clocks and power cap recorded); `e2e` goes through the host-buffer call (H2D of the feature maps + the step + D2H of the
detections).  N > 1: weak scaling, each rank owns its own 64 images per step; the detections of every step reach every
rank inside the timed region.
At N = 1 the line also carries: `other_configs` (C1, C3, C5 and fp32 C2 of BASELINE.json), `library_baseline` (the
reference's own torch ops on CUDA tensors on the same GPU: cuDNN conv + torchvision nms), `nms_latency_bs1` (GPU and CPU
p50) and `cpu_baseline` (the reference's CPU path on the host cores).
`--impl reference` times the reference's own CPU path (oracle/ref_port.py, the torch-CPU port; the reference is pure
Python and /root/reference does not travel to the GPU box) on the host cores, without importing the product package.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

COCO_ANCHORS = [[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]]
TINY_ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
CH = (256, 512, 1024)
SHAPES = [(80, 80), (40, 40), (20, 20)]
STRIDES = [8.0, 16.0, 32.0]
NC = 80
CONF, IOU = 0.25, 0.45
INPUT_SHAPE, IMAGE_SHAPE = (640, 640), (512, 773)
ELEMS_PER_IMG = 2867200                     # feature-map elements per 640x640 image
BYTES_PER_IMG = {"bf16": ELEMS_PER_IMG * 2 + 25200 * 85 * 4, "fp32": ELEMS_PER_IMG * 4 + 25200 * 85 * 4}  # S1: maps in + z out
BYTES_PER_IMG_FUSED = {"bf16": ELEMS_PER_IMG * 2, "fp32": ELEMS_PER_IMG * 4}                              # S3: maps in (+28 B/candidate)
FLOPS_PER_IMG = 2 * 255 * ELEMS_PER_IMG


# ---- parameters: plain torch, shared by both arms (the reference arm never imports the product package) ------------
def make_params(seed=0, nc=NC, ch=CH, anchors=COCO_ANCHORS, im_ref_init=False, obj_bias=-5.0, cls_bias=-3.0):
    """Seeded 'trained-like' IDetect parameters: box rows N(0,.02) as Model.initial_weights
    (nets/yolo.py:120); obj/cls rows scaled so that logits are ~N(obj_bias,1.5) / N(cls_bias,1.5); ia ~ N(0,.02);
    im ~ N(1,.02), or the reference's own initialisation N(0,.02) (nets/common.py:430) with im_ref_init."""
    import torch
    g = torch.Generator().manual_seed(seed)
    na, no = len(anchors[0]) // 2, nc + 5
    p = {"w": [], "b": [], "ia": [], "im": [], "anchors": torch.tensor(anchors).float().view(len(anchors), -1, 2).numpy()}
    for k in ch:
        w = torch.randn(na * no, k, generator=g) * 0.02
        wv = w.view(na, no, k)
        wv[:, 4:, :] = torch.randn(na, no - 4, k, generator=g) * (1.5 / k ** 0.5)
        b = torch.zeros(na, no)
        b[:, 4], b[:, 5:] = obj_bias, cls_bias
        p["w"].append(w)
        p["b"].append(b.view(-1).clone())
        p["ia"].append(torch.randn(k, generator=g) * 0.02)
        im = torch.randn(na * no, generator=g) * 0.02
        p["im"].append(im if im_ref_init else 1.0 + im)
    return p


def make_head(params=None, cls=None, nc=NC, ch=CH, anchors=COCO_ANCHORS):
    """The product's drop-in head module loaded with `params` (B200 arm only)."""
    import torch
    from yolo_continuous_b200.nets import IDetect
    params = params or make_params(nc=nc, ch=ch, anchors=anchors)
    head = (cls or IDetect)(nc, anchors, ch).eval()
    with torch.no_grad():
        for i in range(len(anchors)):
            head.m[i].weight.copy_(params["w"][i].view_as(head.m[i].weight))
            head.m[i].bias.copy_(params["b"][i])
            head.ia[i].implicit.copy_(params["ia"][i].view_as(head.ia[i].implicit))
            head.im[i].implicit.copy_(params["im"][i].view_as(head.im[i].implicit))
    head.stride = torch.tensor(STRIDES)
    return head


def make_maps(bs, seed, dtype, device, ch=CH, shapes=SHAPES):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    return [torch.randn(bs, c, h, w, generator=g, device=device).to(dtype) for c, (h, w) in zip(ch, shapes)]


def time_cpu_port(params, bs, iters, warmup, seed=1234, min_seconds=0.0, conf=CONF, iou=IOU):
    """The reference's CPU path (torch-CPU port) on `bs` images of the workload; returns img/s, cores, median
    seconds per pass, passes.  At least `iters` timed passes, and as many as it takes to fill `min_seconds`."""
    import torch
    from oracle import ref_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    xs = [x.float() for x in make_maps(bs, seed, torch.bfloat16, "cpu")]
    ts = []
    with torch.no_grad():
        it = 0
        while it < warmup + iters or sum(ts) < min_seconds:
            t0 = time.perf_counter()
            ref_port.post_backbone(params, [x.clone() for x in xs], STRIDES, NC, INPUT_SHAPE, IMAGE_SHAPE, True, conf, iou)
            if it >= warmup:
                ts.append(time.perf_counter() - t0)
            it += 1
    return bs / statistics.median(ts), cores, statistics.median(ts), len(ts)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self, wait_s=8.0):
        """Starts nvidia-smi and waits for its first sample (it can take seconds to come up on an 8-GPU box)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < wait_s:
                time.sleep(0.02)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        return len(self.rows)

    def summary(self, lo=0, hi=None):
        rows = self.rows[lo:hi] or self.rows[-1:]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.1)
        self.proc.terminate()


def workload_config(bs, dtype):
    return {"workload": "C2: yolov7 COCO IDetect head (nc=80, 3 anchors x strides 8/16/32, ch 256/512/1024) decode+NMS, "
                        "640x640, synthetic feature maps", "batch_per_gpu": bs, "rows_per_image": 25200,
            "conf_thres": CONF, "nms_thres": IOU, "feature_dtype": dtype,
            "l2": "inputs larger than L2 (feature maps %.0f MB per step)" % (bs * ELEMS_PER_IMG * (2 if dtype == "bf16" else 4) / 1e6)}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  Plain torch + oracle/ only:
    no product module (and so none of its shared objects) is loaded in this process."""
    if rank != 0:
        return
    params = make_params()
    bs = args.bs   # every step processes the full batch of the workload
    ips, cores, sec, _ = time_cpu_port(params, bs, args.steps, args.warmup)
    assert "yolo_continuous_b200" not in sys.modules
    line = {"impl": "reference", "metric": "post_backbone_images_per_sec", "value": ips, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(bs, args.dtype), "images_per_step": bs,
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"all {bs} images of the C2 batch per step (the same bf16-representable feature maps, "
                                       f"upcast to float32 as the reference computes), torch CPU ops (the reference's own "
                                       f"operators, oracle/ref_port.py), {cores} threads, median step"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def pipelining_note(overlap):
    return ("two streams: the NMS kernels of step i run next to the head kernel of step i+1 (double-buffered workspaces); "
            "one CUDA graph launch per step; all K steps complete inside the timed region") if overlap else "single stream"


# ---- helpers of the B200 arm ---------------------------------------------------------------------------------
def timed_gpu(fn, n, warm=3, finish=None):
    """Mean milliseconds per call of fn over n calls between two CUDA events on the current stream."""
    import torch
    for _ in range(warm):
        fn()
    if finish:
        finish()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    if finish:
        finish()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except (OSError, ValueError):
        return {}, "fallback"


def traffic_of(key):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(key)
    except (OSError, ValueError):
        return None


def other_configs(dev, peaks):
    """The other configurations of BASELINE.json on one GPU, device-resident, each with images/s and the fraction of the
    roofline that bounds it (HBM unless stated).  They are parity-test cases first (tests/), measured here so that the
    driver sees them."""
    import torch
    from yolo_continuous_b200.nets import IAuxDetect, IBin
    from yolo_continuous_b200.pipeline import PostBackbone
    hbm = float(peaks.get("hbm_gbs", 6650.0)) * 1e9
    tpeak = float(peaks.get("bf16_tflops", 1624.0)) * 1e12
    out = {}

    def pipelined(pipe, xs, n, warm):
        ms = timed_gpu(lambda: pipe.submit(xs), n, warm, finish=pipe.drain)
        r = pipe.drain() or pipe._views(pipe.cur)
        torch.cuda.synchronize()
        return ms, int(r[3][-1])

    # C1: yolov7-tiny head (cfg/net/yolov7-tiny.yaml: ch 128/256/512, nc = 1), one 640x640 image, conf / iou 0.3
    # (detect.py:271-272).  A single image is latency bound: ms per image is the figure, the HBM fraction is reported
    # only to show that.
    ch1 = (128, 256, 512)
    p1 = make_params(seed=0, nc=1, ch=ch1, anchors=TINY_ANCHORS, obj_bias=-4.5, cls_bias=2.0)   # ~10^2 candidates (195)
    h1 = make_head(p1, nc=1, ch=ch1, anchors=TINY_ANCHORS).to(dev)
    for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "fp32")):
        xs = make_maps(1, 5, dt, dev, ch1)
        pipe = PostBackbone(h1, 1, SHAPES, dt, INPUT_SHAPE, IMAGE_SHAPE, True, 0.3, 0.3, dev, use_graph=True)
        ms = timed_gpu(lambda: pipe.run_device(xs), 50, 5)
        byts = sum(c * h * w for c, (h, w) in zip(ch1, SHAPES)) * (2 if dt == torch.bfloat16 else 4) + \
            (0 if pipe.fused else 25200 * 6 * 4 * 2)
        out[f"c1_tiny_nc1_bs1_{name}"] = {
            "config": "C1: yolov7-tiny head nc=1, 640x640, batch 1, conf 0.3 / iou 0.3, whole step as one CUDA graph",
            "ms_per_image": ms, "images_per_s": 1e3 / ms, "detections": int(pipe.meta[1:][-1]),
            "path": "fused tcgen05 step" if pipe.fused else "head (fp32 maps) + threshold + NMS kernels",
            "roofline": {"bound": "latency (one image: 0.07 GFLOP, 2.9 MB)", "hbm_frac": byts / (ms / 1e3) / hbm}}
        del pipe
    # C3: mAP-eval stress, conf 0.001 / iou 0.65, batch 256; trained-like weights (about 20 k candidates per image spread
    # over 80 classes) and the reference's own ImplicitM initialisation (all logits ~ 0: every one of the 25 200 rows
    # of every image is a candidate).  NMS-bound: the roofline figure is the head kernel's share of the step.
    bs3 = 256
    for name, ref_init in (("c3_map_eval_bs256_trained_like", False), ("c3_map_eval_bs256_ref_init_im", True)):
        head = make_head(make_params(im_ref_init=ref_init)).to(dev)
        xs = make_maps(bs3, 1234, torch.bfloat16, dev)
        pipe = PostBackbone(head, bs3, SHAPES, torch.bfloat16, INPUT_SHAPE, IMAGE_SHAPE, True, 0.001, 0.65, dev,
                            use_graph=False, overlap=True)
        ms, det = pipelined(pipe, xs, 6, 2)
        ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(4)]
        for e in ev:
            pipe.run_device(xs, head_events=e)
        pipe.wait()
        torch.cuda.synchronize()
        hm = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
        nm = statistics.mean(e[2].elapsed_time(e[3]) for e in ev)
        out[name] = {"config": "C3: conf 0.001 / iou 0.65, 25 200 rows x 80 classes per image, batch 256, bf16 maps, "
                               "pipelined fused step", "ms_per_step": ms, "images_per_s": bs3 / ms * 1e3,
                     "detections_per_step": det, "head_kernel_ms": hm, "nms_kernels_ms": nm,
                     "roofline": {"bound": "NMS: SM throughput / latency (sum n_c^2/2 IoU tests)",
                                  "head_tensor_frac": FLOPS_PER_IMG * bs3 / (hm / 1e3) / tpeak,
                                  "step_hbm_frac": (BYTES_PER_IMG_FUSED["bf16"] * bs3 + 28 * det) / (ms / 1e3) / hbm}}
        del pipe, xs
        torch.cuda.empty_cache()
    # C5: IAuxDetect and IBin at 1280x1280 (160/80/40 maps: 100 800 rows per image), 16 images per GPU (128 over 8)
    shapes5 = [(160, 160), (80, 80), (40, 40)]
    bs5 = 16
    elems5 = sum(c * h * w for c, (h, w) in zip(CH, shapes5))
    for name, cls, ch in (("c5_iauxdetect_1280_bs16", IAuxDetect, CH * 2), ("c5_ibin_1280_bs16", IBin, CH)):
        head = cls(NC, COCO_ANCHORS, ch).eval()
        with torch.no_grad():   # trained-like objectness / class biases (as make_params): ~10^2 candidates per image
            for im in head.im:   # ImplicitM ~ N(1, .02) as make_params (the reference's N(0, .02) scales every logit to ~0)
                im.implicit.copy_(1.0 + 0.02 * torch.randn(im.implicit.shape, generator=torch.Generator().manual_seed(11)))
            g5 = torch.Generator().manual_seed(12)
            for conv in head.m:      # objectness / class logits ~ N(-5, 1.5) / N(-3, 1.5), as make_params
                k = conv.weight.shape[1]
                w = conv.weight.view(head.na, head.no, k)
                w[:, head.no - NC - 1:, :] = torch.randn(head.na, NC + 1, k, generator=g5) * (1.5 / k ** 0.5)
                b = conv.bias.view(head.na, head.no)
                b[:, head.no - NC - 1] = -5.0
                b[:, head.no - NC:] = -3.0
        head = head.to(dev)
        head.stride = torch.tensor(STRIDES)
        head.return_raw = False
        head.compute_aux_in_eval = False   # the reference's dead aux convolution in eval (nets/iaux_detect.py:37-38)
        xs = make_maps(bs5, 7, torch.bfloat16, dev, ch, shapes5 * (len(ch) // 3))
        with torch.no_grad():
            ms = timed_gpu(lambda: head(list(xs)), 10, 2)
            z = head(list(xs))[0]
        byts = (elems5 * 2 + z.shape[1] * z.shape[2] * 4) * bs5
        flops = 2 * head.na * head.no * elems5 * bs5
        # the whole step (head -> NMS, z never written) through the pipelined fused path, C2 thresholds
        head.return_raw = True
        pipe = PostBackbone(head, bs5, shapes5, torch.bfloat16, (1280, 1280), (720, 1280), True, CONF, IOU, dev,
                            use_graph=False, overlap=True)
        ms_step, det = pipelined(pipe, xs[:3], 10, 3)
        fused_step = pipe.fused
        del pipe
        out[name] = {"config": f"C5: {cls.__name__} at 1280x1280, batch 16, bf16 maps: forward -> z (z only; "
                               f"{'aux convolution skipped, ' if cls is IAuxDetect else ''}module call incl. allocation) and "
                               f"the fused step head -> NMS (conf {CONF} / iou {IOU}, pipelined graphs)",
                     "forward_ms": ms, "images_per_s": bs5 / ms * 1e3,
                     "step_ms": ms_step, "step_images_per_s": bs5 / ms_step * 1e3, "step_detections": det, "step_fused": fused_step,
                     "step_roofline": {"bound": "hbm", "frac": elems5 * 2 * bs5 / (ms_step / 1e3) / hbm,
                                       "tensor_frac": 2 * head.na * head.no * elems5 * bs5 / (ms_step / 1e3) / tpeak},
                     "roofline": {"bound": "hbm", "achieved": byts / (ms / 1e3) / 1e9, "peak": hbm / 1e9, "unit": "GB/s",
                                  "frac": byts / (ms / 1e3) / hbm, "tensor_tflops": flops / (ms / 1e3) / 1e12}}
        del xs, z
        torch.cuda.empty_cache()
    # C2 through the drop-in forward that materialises z (what IDetect.forward returns to a reference user): S1, bf16 maps,
    # z only and with the raw maps list (the reference's return value, twice the output bytes)
    head = make_head().to(dev)
    xs = make_maps(64, 1234, torch.bfloat16, dev)
    s1 = {}
    for name, raw in (("z_only", False), ("z_and_raw_maps", True)):
        head.return_raw = raw
        with torch.no_grad():
            ms_f = timed_gpu(lambda: head(list(xs)), 20, 3)
        byts = (BYTES_PER_IMG["bf16"] + (25200 * 85 * 4 if raw else 0)) * 64
        s1[name] = {"forward_ms": ms_f, "images_per_s": 64 / ms_f * 1e3,
                    "roofline": {"bound": "hbm", "achieved": byts / (ms_f / 1e3) / 1e9, "peak": hbm / 1e9, "unit": "GB/s",
                                 "frac": byts / (ms_f / 1e3) / hbm, "bytes": byts}}
    out["c2_forward_z_bf16_bs64"] = dict(s1, config="C2, IDetect.forward -> z [64, 25200, 85] float32 from bf16 maps (S1: 14.30 MB/img; "
                                         "with the raw maps 22.87 MB/img), module call incl. allocation; head_tc2_kernel, rows by halves",
                                         kernel="head_tc2_kernel<FUSED=false> (CTA pairs, z rows by halves through 16-row slabs)")
    del xs
    torch.cuda.empty_cache()
    # C2 with float32 feature maps (the reference's own precision, 1e-5 parity): drop-in forward and head -> NMS
    head = make_head().to(dev)
    xs = make_maps(64, 1234, torch.float32, dev)
    head.return_raw = False
    with torch.no_grad():
        ms_f = timed_gpu(lambda: head(list(xs)), 10, 2)
    head.return_raw = True
    pipe = PostBackbone(head, 64, SHAPES, torch.float32, INPUT_SHAPE, IMAGE_SHAPE, True, CONF, IOU, dev, use_graph=True)
    ms_s = timed_gpu(lambda: pipe.run_device(xs), 10, 2)
    out["c2_fp32_maps_bs64"] = {
        "config": "C2 with float32 feature maps (1e-5 parity mode): IDetect.forward -> z, and the whole step (head + "
                  "threshold + NMS) as one CUDA graph",
        "forward_ms": ms_f, "forward_images_per_s": 64 / ms_f * 1e3, "step_ms": ms_s, "images_per_s": 64 / ms_s * 1e3,
        "detections_per_step": int(pipe.meta[64:][-1]),
        "roofline": {"bound": "hbm", "achieved": BYTES_PER_IMG["fp32"] * 64 / (ms_f / 1e3) / 1e9, "peak": hbm / 1e9,
                     "unit": "GB/s", "frac": BYTES_PER_IMG["fp32"] * 64 / (ms_f / 1e3) / hbm, "traffic": traffic_of("head_fp32_bs64"),
                     "kernel": "head_tcs_kernel (float32 maps split into fp16 hi/lo in the kernel, 3 tcgen05 MMAs per k-step)",
                     "tensor_tflops": FLOPS_PER_IMG * 64 / (ms_f / 1e3) / 1e12,
                     "note": "S1 fp32: 20.04 MB/img (maps in + z out), ceiling 327 k img/s"}}
    return out


def library_baseline(params, dev, bs=16):
    """The bar SURVEY.md section 2.3 sets: the reference's own torch code on CUDA tensors on the same B200 -- cuDNN/cuBLAS
    1x1 convolution + ~50 elementwise launches per forward (nets/idetect.py:26-45), then the per-image / per-class Python
    loop around torchvision's sm_100 nms kernels with its host round trips (detect.py:90-144) -- with TF32 convolutions
    off (the parity setting) and on (torch's default)."""
    import torch
    from oracle import ref_port
    pd = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v) for k, v in params.items()}
    xs = [x.float() for x in make_maps(bs, 1234, torch.bfloat16, dev)]
    res = {"images_per_pass": bs, "what": "oracle/ref_port.py (the reference's torch ops, unfused) on CUDA tensors: "
                                          "torch conv2d (cuDNN) + elementwise decode + torchvision.ops.nms per class"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (("tf32_off", False), ("tf32_on", True)):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
            ts, hs = [], []
            with torch.no_grad():
                for it in range(6):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    z, _ = ref_port.idetect_forward(pd, [x.clone() for x in xs], STRIDES)
                    torch.cuda.synchronize()
                    t1 = time.perf_counter()
                    z[..., 0] /= INPUT_SHAPE[1]; z[..., 2] /= INPUT_SHAPE[1]
                    z[..., 1] /= INPUT_SHAPE[0]; z[..., 3] /= INPUT_SHAPE[0]
                    out = ref_port.non_max_suppression(z, NC, INPUT_SHAPE, IMAGE_SHAPE, True, CONF, IOU)
                    torch.cuda.synchronize()
                    t2 = time.perf_counter()
                    if it >= 2:
                        hs.append(t1 - t0)
                        ts.append(t2 - t0)
            res[name] = {"images_per_s": bs / statistics.median(ts), "head_forward_images_per_s": bs / statistics.median(hs),
                         "nms_ms_per_image": (statistics.median(ts) - statistics.median(hs)) / bs * 1e3,
                         "detections": sum(0 if o is None else len(o) for o in out)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return res


def nms_latency(head, xs, dev, tdt):
    """NMS latency at batch 1 (second metric of BASELINE.json): threshold/compaction + per-class NMS + letterbox undo on
    one decoded 640x640 image (z resident in HBM), preallocated buffers, replayed as a CUDA graph; beside it the
    reference's CPU NMS (oracle/ref_port.py = detect.py:90-144 with torchvision.ops.nms) on the same image."""
    import torch
    from oracle import ref_port
    from yolo_continuous_b200 import _lib
    from yolo_continuous_b200.pipeline import PostBackbone
    res = {}
    zsrc = PostBackbone(head, 8, SHAPES, tdt, INPUT_SHAPE, IMAGE_SHAPE, True, CONF, IOU, dev, use_graph=False, fused=False)
    zsrc.run_device([x[:8].contiguous() for x in xs])
    torch.cuda.synchronize()
    z_img = zsrc.z[3:4].clone()
    z_cpu = z_img.cpu()
    z_cpu[..., 0] /= INPUT_SHAPE[1]; z_cpu[..., 2] /= INPUT_SHAPE[1]
    z_cpu[..., 1] /= INPUT_SHAPE[0]; z_cpu[..., 3] /= INPUT_SHAPE[0]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for name, conf, iou in (("c2_conf0.25_iou0.45", CONF, IOU), ("c3_conf0.001_iou0.65", 0.001, 0.65)):
        p1 = PostBackbone(head, 1, SHAPES, tdt, INPUT_SHAPE, IMAGE_SHAPE, True, conf, iou, dev, use_graph=False, fused=False)
        p1.z.copy_(z_img)

        def nms_only():
            m1 = p1.meta.data_ptr()
            _lib.check(_lib.lib.yc_nms_batched(p1.z.data_ptr(), p1.nms_params, p1.ws.data_ptr(), p1.ws.numel(),
                                               p1.out_rows.data_ptr(), p1.out_idx.data_ptr(), m1, m1 + 4,
                                               _lib.stream_ptr(dev)), "yc_nms_batched")
        nms_only()
        torch.cuda.synchronize()
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            nms_only()
        lat = []
        for i in range(110):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g1.replay()
            b.record()
            b.synchronize()
            if i >= 10:
                lat.append(a.elapsed_time(b))
        cpu = []
        n_cpu = 0
        with torch.no_grad():
            for i in range(9 if conf > 0.01 else 5):
                zc = z_cpu.clone()
                t0 = time.perf_counter()
                o = ref_port.non_max_suppression(zc, NC, INPUT_SHAPE, IMAGE_SHAPE, True, conf, iou)
                cpu.append((time.perf_counter() - t0) * 1e3)
                n_cpu = 0 if o[0] is None else len(o[0])
        res[name] = {"p50_ms": statistics.median(lat), "detections": int(p1.meta[0].item()),
                     "cpu_p50_ms": statistics.median(cpu[2:]), "cpu_cores": cores, "cpu_detections": n_cpu,
                     "cpu_kind": "port (detect.py:90-144 restated over torch CPU ops + torchvision.ops.nms)"}
        del p1
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bs", type=int, default=64)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs / library_baseline / sustained loop")
    ap.add_argument("--sustain-seconds", type=float, default=2.5)
    ap.add_argument("--unfused", action="store_true", help="materialise z (drop-in forward) and run the stand-alone threshold kernel")
    ap.add_argument("--no-overlap", action="store_true", help="NMS kernels on the head kernel's stream (no second stream)")
    ap.add_argument("--profile", action="store_true", help="device-resident loop only (for ncu runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from yolo_continuous_b200 import _lib
    from yolo_continuous_b200.parallel import DetectionGather
    from yolo_continuous_b200.pipeline import PostBackbone

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.check(_lib.lib.yc_device_check(local), "yc_device_check")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("YC_NCCL_CTAS", "") not in ("", "0"):
            os.environ["NCCL_MAX_CTAS"] = os.environ["NCCL_MIN_CTAS"] = os.environ["YC_NCCL_CTAS"]
        dist.init_process_group("nccl", device_id=dev)
    if os.environ.get("YC_RESERVE_SMS"):
        _lib.lib.yc_reserve_sms(int(os.environ["YC_RESERVE_SMS"]))
    W = max(args.warmup, 3)
    K = args.steps
    tdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    overlap = not args.no_overlap and not args.unfused and args.dtype == "bf16"
    params = make_params()
    head = make_head(params).to(dev)
    pipe = PostBackbone(head, args.bs, SHAPES, tdt, INPUT_SHAPE, IMAGE_SHAPE, True, CONF, IOU, dev, use_graph=False,
                        fused=not args.unfused, double_buffer=world > 1, overlap=overlap)
    # rows per rank in the fixed-size exchange (C2 produces ~2.7k per 64 images; a rank with more says so in its
    # header and the remainder is fetched with gather_detections): 115 KB, ~45 us with one NCCL CTA
    GATHER_ROWS = int(os.environ.get("YC_GATHER_ROWS", "4096"))
    # one collective per GATHER_EVERY steps (see DetectionGather): 16 steps = 1024 images per rank per exchange
    # (the host side of a NCCL call costs 0.2-0.6 ms at 2-8 ranks: one per step would make the loop host-bound)
    GATHER_EVERY = int(os.environ.get("YC_GATHER_EVERY", "16"))
    # Exchange of the detections (world > 1).  Default: the repo's own kernels over NVLink peer memory (PeerExchange: one
    # push kernel + one wait kernel per step inside the step's CUDA graph, no NCCL call and no host work per step);
    # YC_EXCHANGE=nccl (or a box without CUDA IPC peer access) selects the grouped NCCL all-gather.
    gather, xchg = None, None
    if world > 1:
        if os.environ.get("YC_EXCHANGE", "peer") == "peer" and overlap:
            try:
                from yolo_continuous_b200.parallel import PeerExchange
                xchg = PeerExchange(pipe.hdr_ints, args.bs, GATHER_ROWS, dev)
                pipe.attach_exchange(xchg)
            except _lib.YcError as e:   # every rank fails or succeeds alike (same box, same driver)
                print(f"[bench] peer exchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
                xchg = None
        if xchg is None:
            gather = DetectionGather(pipe.message(GATHER_ROWS).numel(), dev, every=GATHER_EVERY)
    xs = make_maps(args.bs, 1234 + rank, tdt, dev)
    for d_, h_ in zip(xs, pipe.x_host):
        h_.copy_(d_)

    if overlap:
        # the pipeline runs on its own non-blocking stream with the higher priority (the NMS kernels of the previous
        # batch are on the default-priority tail stream); measured neutral on the step time, kept so that the legacy
        # default stream's implicit synchronisation never enters the picture
        torch.cuda.synchronize()
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))

    pipelined = overlap   # one CUDA graph per step: head of batch i next to the NMS kernels of batch i-1

    def step():
        if pipelined:
            prev = pipe.submit(xs)           # with the peer exchange attached the graph also pushes / awaits the messages
            if gather is not None and prev is not None:
                gather.gather_async(pipe.message(GATHER_ROWS, previous=True))
        else:
            pipe.run_device(xs)
            if gather is not None:
                gather.gather_async(pipe.message(GATHER_ROWS), stream=pipe.tail_stream)

    def finish():
        slot = None
        if pipelined:
            pipe.drain()
            if gather is not None:
                slot = gather.gather_async(pipe.message(GATHER_ROWS))
        if xchg is not None:
            pipe.wait()
            xchg.wait_all()                  # the messages of all steps of all ranks have arrived
        if gather is not None:
            s2 = gather.flush()
            slot = s2 if s2 is not None else slot
            gather.wait()
        pipe.wait()
        return slot

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput: K steps between two events (all work of the K steps completes inside) ----------
    for _ in range(W):
        step()
    slot = finish()
    exchange_check = None
    if world > 1:
        # once, outside the timed region: what the exchange delivered == what the ranks produced.  The last group holds
        # this rank's latest message(s): its own slot must be bit-identical to its local result, and every rank's
        # header must agree with a plain all-gather of the local counts.
        torch.cuda.synchronize()
        rows_l, _, counts_l, offs_l = pipe._views(pipe.cur)
        total_l = int(offs_l[-1])
        if xchg is not None:
            sp, sw, err = xchg.state()
            assert os.environ.get("YC_XCHG_DEBUG") or (err == 0 and sp == sw == W), f"peer exchange state {(sp, sw, err)} after {W} steps"
            got = xchg.unpack(sw - 1)
            last = got[rank]
        else:
            n_in_group = (W - 1) % GATHER_EVERY + 1
            got = gather.unpack(slot, args.bs, pipe.hdr_ints, GATHER_ROWS, n=n_in_group)
            got = [g_[-1] if GATHER_EVERY > 1 else g_ for g_ in got]
            last = got[rank]
        ok = torch.equal(last[0], counts_l) and int(last[1]) == total_l and \
            torch.equal(last[2][:min(total_l, GATHER_ROWS)], rows_l[:min(total_l, GATHER_ROWS)])
        all_counts = [torch.empty_like(counts_l) for _ in range(world)]
        dist.all_gather(all_counts, counts_l.contiguous())
        for r in range(world):
            ok = ok and torch.equal(got[r][0], all_counts[r])
        assert ok or os.environ.get("YC_XCHG_DEBUG"), "detection exchange: gathered messages differ from what the ranks produced"
        exchange_check = "gathered == produced (own rows bit-identical; all ranks' counts vs a plain all-gather)"
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    m0 = sampler.mark()
    t0.record()
    host_t0 = time.perf_counter()
    for i in range(K):
        step()
    host_loop = time.perf_counter() - host_t0
    finish()
    t1.record()
    sync_all()
    ms = t0.elapsed_time(t1)
    if world > 1:   # max over ranks (also: every rank must derive the same step counts from it below)
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # ---- the same K steps issued call by call (two streams, no graph) with the head kernel and the NMS kernels
    # bracketed by CUDA events on the streams they are launched on: per-kernel durations for the roofline ----------
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(K)]   # head start/end, NMS start/end
    for _ in range(3):
        pipe.run_device(xs)
    pipe.wait()
    sync_all()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for i in range(K):
        pipe.run_device(xs, head_events=ev[i])
    pipe.wait()
    u1.record()
    sync_all()
    eager_ms = u0.elapsed_time(u1)
    # single-stream pass (no overlap): what a serialised ncu launch list sees
    serial_share = None
    if overlap:
        sp = PostBackbone(head, args.bs, SHAPES, tdt, INPUT_SHAPE, IMAGE_SHAPE, True, CONF, IOU, dev, use_graph=False)
        sev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(20)]
        for _ in range(3):
            sp.run_device(xs)
        for i in range(20):
            sp.run_device(xs, head_events=sev[i])
        torch.cuda.synchronize()
        sh_, st_ = (statistics.mean(e[0].elapsed_time(e[1]) for e in sev), statistics.mean(e[2].elapsed_time(e[3]) for e in sev))
        serial_share = {"head_ms": sh_, "nms_kernels_ms": st_, "share": sh_ / (sh_ + st_)}
        del sp
    # ---- the same step for >= 2 s: what the GPU sustains (clocks, power cap) ---------------------------------------
    sustained = None
    if not args.no_extras and not args.profile:
        n_s = max(K, int(args.sustain_seconds * 1e3 / (ms / K)))
        n_s = (n_s + GATHER_EVERY - 1) // GATHER_EVERY * GATHER_EVERY
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        ms0 = sampler.mark()
        s0.record()
        for i in range(n_s):
            step()
        finish()
        s1.record()
        sync_all()
        ms1 = sampler.mark()
        s_ms = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([s_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_ms = float(t.item())
        sustained = {"steps": n_s, "seconds": s_ms / 1e3, "ms_per_step": s_ms / n_s,
                     "value": world * args.bs * n_s / (s_ms / 1e3), "unit": "images/s",
                     "clocks": sampler.summary(ms0, ms1) if rank == 0 else None}
    m1 = sampler.mark()
    clocks = sampler.summary(m0, m1) if rank == 0 else None   # timed region + per-kernel passes + sustained loop
    sampler.stop()
    head_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    tail_ms = statistics.mean(e[2].elapsed_time(e[3]) for e in ev)
    value = world * args.bs * K / (ms / 1e3)
    counts_host = pipe.meta[:args.bs].cpu().numpy()
    n_det = int(counts_host.sum())

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "value": value, "ms_per_step": ms / K, "head_ms": head_ms, "nms_ms": tail_ms,
                              "eager_ms_per_step": eager_ms / K, "host_enqueue_ms_per_step": host_loop / K * 1e3}))
        return

    # ---- end to end through the host-buffer call ------------------------------------------------------------
    # every step: H2D of that step's feature maps from pinned host memory, the kernels, D2H of its detections; with the
    # pipelined call the H2D of step i+1 runs under the kernels and the read-back of step i
    def host_step():
        if overlap:
            return pipe.submit_host()
        out_ = pipe.run_host()
        return out_

    for _ in range(3):
        host_step()
    if overlap:
        pipe.drain_host()
    sync_all()
    e0 = time.perf_counter()
    total_rows, out = 0, None
    for _ in range(K):
        o = host_step()
        out = o if o is not None else out
        if gather is not None:
            gather.gather_async(pipe.message(GATHER_ROWS), stream=pipe.tail_stream)
    if overlap:
        out = pipe.drain_host()
    if xchg is not None:
        xchg.wait_all(pipe.tail_stream)
    if gather is not None:
        gather.flush()
        gather.wait()
    torch.cuda.synchronize()
    e_ms = (time.perf_counter() - e0) * 1e3
    total_rows = sum(0 if o is None else len(o) for o in out)
    if world > 1:
        t = torch.tensor([e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
    e2e_value = world * args.bs * K / (e_ms / 1e3)
    # the box's concurrent host->device ceiling at this rank count: the same bytes per step as plain pinned copies on
    # every rank at once (no kernels), so that e2e can be read as a fraction of what the PCIe fabric gives N ranks
    copy_stream = torch.cuda.Stream(device=dev)
    sync_all()
    c0 = time.perf_counter()
    n_copy = max(4, min(K, 20))
    with torch.cuda.stream(copy_stream):
        for _ in range(n_copy):
            for d_, h_ in zip(pipe.x_dev, pipe.x_host):
                d_.copy_(h_, non_blocking=True)
    copy_stream.synchronize()
    c_s = time.perf_counter() - c0
    if world > 1:
        t = torch.tensor([c_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c_s = float(t.item())
    h2d_ceiling_gbs = world * n_copy * pipe.h2d_bytes() / c_s / 1e9
    h2d_ceiling_ips = world * n_copy * args.bs / c_s

    nms_lat = nms_latency(head, xs, dev, tdt) if rank == 0 and not args.no_extras else None
    if xchg is not None:
        sp, sw, err = xchg.state()
        assert os.environ.get("YC_XCHG_DEBUG") or (err == 0 and sp == sw), f"peer exchange state {(sp, sw, err)} at the end of the run"
        pipe.attach_exchange(None)
        xchg.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_launch = (BYTES_PER_IMG_FUSED if pipe.fused else BYTES_PER_IMG)[args.dtype] * args.bs
    # The roofline figure is the head kernel timed ALONE (the single-stream pass: nothing else runs on the GPU while it
    # does, which is also what the ncu capture sees); next to the NMS kernels of the previous batch (the pipelined step)
    # the same launch takes ~10 % longer -- that duration is kept as kernel_ms_beside_nms and bounds the step.
    head_beside_ms = head_ms
    if serial_share:
        head_ms = serial_share["head_ms"]
    achieved = bytes_launch / (head_ms / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"head_{'fused_' if pipe.fused else ''}{args.dtype}_bs{args.bs}")
    except (OSError, ValueError):
        pass
    tflops = FLOPS_PER_IMG * args.bs / (head_ms / 1e3) / 1e12
    kname = "head_generic_kernel" if args.dtype != "bf16" else \
        ("head_tc2_kernel (CTA pairs, cta_group::2)" if pipe.fused and os.environ.get("YC_TC_2CTA", "1") != "0" else "head_tc_kernel")
    roofline = {"kernel": kname,
                "stage": "S3 fused head->candidates (z never written)" if pipe.fused else "S1 head->z",
                "peak_source": peak_src, "traffic": traffic, "kernel_ms": head_ms,
                "kernel_ms_beside_nms": head_beside_ms,
                "kernel_share_of_step": head_beside_ms / (ms / K),
                "kernel_timing": "kernel_ms: CUDA events around each of 20 launches in a single-stream pass (the kernel alone "
                                 "on the GPU, the NMS kernels after it); kernel_ms_beside_nms: the same around each launch of "
                                 f"a second pass of the {K} steps issued call by call on two streams ({eager_ms / K:.4f} ms per "
                                 "step), where the NMS kernels of the previous batch share the SMs with it, as in the "
                                 "graph-replayed timed region (which cannot hold per-kernel events)",
                # the three NMS kernels of a step, timed on their own stream (they run next to the NEXT step's head
                # kernel when pipelined, which stretches them); kernel_share_serialised comes from a single-stream
                # pass of 20 steps and is what an ncu launch list (serialised) shows
                "nms_kernels_ms": tail_ms,
                "kernel_share_serialised": serial_share if serial_share else head_ms / (head_ms + tail_ms),
                "bytes_per_launch": bytes_launch,
                "flops_per_launch": FLOPS_PER_IMG * args.bs}
    if pipe.fused:
        # S3 moves 5.73 MB/img but still needs 1.462 GFLOP/img: the tensor pipe binds first.
        # The kernel is timed launch by launch inside a short eager pass, i.e. at burst clocks: its fraction is taken
        # against the BURST peak; the sustained loop's whole-step rate is taken against the SUSTAINED peak.
        tpeak = float(peaks.get("bf16_tflops", 1624.0))
        tpeak_s = float(peaks.get("bf16_tflops_sustained", 1400.0))
        roofline.update({"bound": "tensor", "achieved": tflops, "peak": tpeak, "unit": "TFLOP/s", "frac": tflops / tpeak,
                         "peak_kind": "bf16_tflops (burst): the kernel is timed alone, launch by launch",
                         "hbm_gbs": achieved, "hbm_frac": achieved / hbm_peak})
        if sustained:
            st = FLOPS_PER_IMG * args.bs / (sustained["ms_per_step"] / 1e3) / 1e12
            roofline.update({"sustained_step_tflops": st, "sustained_peak": tpeak_s, "frac_sustained": st / tpeak_s,
                             "sustained_kind": "whole step (head kernel back to back, NMS kernels beside it) over the "
                                               "sustained loop vs bf16_tflops_sustained"})
    else:
        roofline.update({"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "tensor_tflops": tflops})
    line = {
        "metric": "post_backbone_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
        "config": dict(workload_config(args.bs, args.dtype), pipelining=pipelining_note(overlap),
                       **({"exchange": (f"own kernels over NVLink peer memory (CUDA IPC): per step one push kernel stores the header + "
                                        f"the first {GATHER_ROWS} detection rows into every rank's receive buffer and raises a "
                                        f"flag, one wait kernel awaits earlier flags -- a third branch of the step's CUDA graph, "
                                        f"next to the head and NMS kernels; no NCCL call, no host work" if xchg is not None else
                                        f"one NCCL all-gather per {GATHER_EVERY} steps: per rank and step the header + the first "
                                        f"{GATHER_ROWS} detection rows") + "; every step's detections reach every rank inside "
                                       "the timed region", "exchange_check": exchange_check} if world > 1 else {})),
        "clocks": clocks,
        "sustained": sustained,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": pipe.h2d_bytes(),
                "d2h_bytes_per_step": pipe.d2h_bytes(total_rows), "ms_per_step": e_ms / K,
                "h2d_ceiling": {"gb_per_s_all_ranks": h2d_ceiling_gbs, "images_per_s": h2d_ceiling_ips,
                                "frac": e2e_value / h2d_ceiling_ips,
                                "what": f"plain pinned cudaMemcpyAsync of the same {pipe.h2d_bytes()} bytes per step on all "
                                        f"{world} rank(s) at once, no kernels"}},
        "gpu_launches": K * (pipe.kernels_per_step + (2 if xchg is not None else 0)),
        "roofline": roofline,
        "detections_per_step": n_det, "nms_latency_bs1": nms_lat,
    }
    if world == 1 and not args.no_extras:
        line["other_configs"] = other_configs(dev, peaks)
        line["library_baseline"] = library_baseline(params, dev)
    if world == 1 and not args.no_cpu_baseline:
        ips, cores, sec, n = time_cpu_port(params, 16, 5, 2, min_seconds=12.0)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"16 images of the same workload per pass, {n} timed passes over "
                                          f"{sec * n:.0f} s (median {sec:.3f} s per pass), torch CPU ops with {cores} threads"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
