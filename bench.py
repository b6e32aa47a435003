#!/usr/bin/env python
"""bench.py -- post-backbone images/sec of the B200 path (and the reference's CPU path beside it).

Workload (BASELINE.json configs[1], "C2"): yolov7 COCO IDetect head (80 classes, 3 anchors x 3 strides,
ch 256/512/1024) at 640x640, batch 64 per GPU, synthetic feature maps, seeded trained-like weights,
conf 0.25 / iou 0.45.  One "step" = one pass of the hot path over one batch:
    head 1x1 conv (+ImplicitA/M) -> sigmoid + box decode -> conf threshold + compaction -> per-class NMS
    -> letterbox undo.
`value` is measured with inputs resident in HBM; `e2e` goes through the host-buffer call
(PostBackbone.run_host: H2D of the feature maps + the step + D2H of the detections).
N > 1: weak scaling, each rank owns its own 64 images; one NCCL all-gather of detections per step.
`--impl reference` times the reference's own CPU path (oracle/ref_port.py, the torch-CPU port; the
reference is pure Python and /root/reference does not travel to the GPU box) on the host cores.
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
CH = (256, 512, 1024)
SHAPES = [(80, 80), (40, 40), (20, 20)]
STRIDES = [8.0, 16.0, 32.0]
NC = 80
CONF, IOU = float(os.environ.get("YC_BENCH_CONF", "0.25")), 0.45   # YC_BENCH_CONF: kernel experiments only
INPUT_SHAPE, IMAGE_SHAPE = (640, 640), (512, 773)
BYTES_PER_IMG = {"bf16": 2867200 * 2 + 25200 * 85 * 4, "fp32": 2867200 * 4 + 25200 * 85 * 4}  # S1: maps in + z out
BYTES_PER_IMG_FUSED = {"bf16": 2867200 * 2, "fp32": 2867200 * 4}                              # S3: maps in (+28 B/candidate)
FLOPS_PER_IMG = 2 * 255 * 2867200


def make_head(seed=0):
    """IDetect with seeded 'trained-like' parameters: box rows N(0,.02) as
    Model.initial_weights (nets/yolo.py:120); obj/cls rows scaled so that logits are ~N(-5,1.5) /
    N(-3,1.5); ia ~ N(0,.02); im ~ N(1,.02)."""
    import torch
    from yolo_continuous_b200.nets import IDetect
    g = torch.Generator().manual_seed(seed)
    head = IDetect(NC, COCO_ANCHORS, CH).eval()
    with torch.no_grad():
        for i, conv in enumerate(head.m):
            k = conv.weight.shape[1]
            w = torch.randn(conv.weight.shape, generator=g) * 0.02
            wv = w.view(head.na, head.no, k)
            wv[:, 4:, :] = torch.randn(head.na, head.no - 4, k, generator=g) * (1.5 / k ** 0.5)
            conv.weight.copy_(w)
            b = torch.zeros(head.na, head.no)
            b[:, 4], b[:, 5:] = -5.0, -3.0
            conv.bias.copy_(b.view(-1))
            head.ia[i].implicit.copy_(torch.randn(head.ia[i].implicit.shape, generator=g) * 0.02)
            head.im[i].implicit.copy_(1.0 + torch.randn(head.im[i].implicit.shape, generator=g) * 0.02)
    head.stride = torch.tensor(STRIDES)
    return head


def port_params(head):
    return {"anchors": head.anchor_grid.detach().cpu().reshape(3, -1, 2).numpy(),
            "w": [m.weight.detach().cpu()[:, :, 0, 0] for m in head.m], "b": [m.bias.detach().cpu() for m in head.m],
            "ia": [a.implicit.detach().cpu().reshape(-1) for a in head.ia],
            "im": [m.implicit.detach().cpu().reshape(-1) for m in head.im]}


def time_cpu_port(head, bs, iters, warmup, seed=1234, min_seconds=0.0):
    """The reference's CPU path (torch-CPU port) on `bs` images of the workload; returns img/s, cores, median
    seconds per pass, passes.  At least `iters` timed passes, and as many as it takes to fill `min_seconds`."""
    import torch
    from oracle import ref_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(bs, c, h, w, generator=g).to(torch.bfloat16).float() for c, (h, w) in zip(CH, SHAPES)]
    p = port_params(head)
    ts = []
    with torch.no_grad():
        it = 0
        while it < warmup + iters or sum(ts) < min_seconds:
            t0 = time.perf_counter()
            ref_port.post_backbone(p, [x.clone() for x in xs], STRIDES, NC, INPUT_SHAPE, IMAGE_SHAPE, True, CONF, IOU)
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
            self.skip = len(self.rows)   # samples taken before the timed region (idle clocks)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.rows = self.rows[getattr(self, "skip", 0):] or self.rows
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    head = make_head()
    bs = 16   # images per step: a bounded sample of the 64-image batch (about 0.1-0.2 s of CPU work per step)
    ips, cores, sec, _ = time_cpu_port(head, bs, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "post_backbone_images_per_sec", "value": ips, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(64, "f32"), pipelining="none (CPU)"), "images_per_step": bs,
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{bs} of the 64 images of the C2 batch per step, torch CPU ops (the reference's "
                                       f"own operators), {cores} threads"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(bs, dtype):
    return {"workload": "C2: yolov7 COCO IDetect head (nc=80, 3 anchors x strides 8/16/32, ch 256/512/1024) decode+NMS, "
                        "640x640, synthetic feature maps", "batch_per_gpu": bs, "rows_per_image": 25200,
            "conf_thres": CONF, "nms_thres": IOU, "feature_dtype": dtype,
            "l2": "inputs larger than L2 (feature maps %.0f MB per step)" % (bs * 2867200 * (2 if dtype == "bf16" else 4) / 1e6)}


def pipelining_note(overlap):
    return ("two streams: the NMS kernels of step i run next to the head kernel of step i+1 (double-buffered workspaces); "
            "one CUDA graph launch per step; all K steps complete inside the timed region") if overlap else "single stream"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bs", type=int, default=64)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        # the one collective of the path is a ~0.5 MB all-gather per step: one NCCL CTA, on the SM the head kernel
        # leaves free for it (yc_reserve_sms), so that it runs next to the head kernel instead of displacing a CTA
        if os.environ.get("YC_NCCL_CTAS", "") not in ("", "0"):
            os.environ["NCCL_MAX_CTAS"] = os.environ["NCCL_MIN_CTAS"] = os.environ["YC_NCCL_CTAS"]
        dist.init_process_group("nccl", device_id=dev)
    if os.environ.get("YC_RESERVE_SMS"):
        _lib.lib.yc_reserve_sms(int(os.environ["YC_RESERVE_SMS"]))
    W = max(args.warmup, 3)
    K = args.steps
    tdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    overlap = not args.no_overlap and not args.unfused and args.dtype == "bf16"
    head = make_head().to(dev)
    pipe = PostBackbone(head, args.bs, SHAPES, tdt, INPUT_SHAPE, IMAGE_SHAPE, True, CONF, IOU, dev, use_graph=False,
                        fused=not args.unfused, double_buffer=world > 1, overlap=overlap)
    # rows per rank in the fixed-size exchange (C2 produces ~2.7k per 64 images; a rank with more says so in its
    # header and the remainder is fetched with gather_detections): 115 KB, ~45 us with one NCCL CTA
    GATHER_ROWS = int(os.environ.get("YC_GATHER_ROWS", "4096"))
    # one collective per GATHER_EVERY steps (see DetectionGather): 16 steps = 1024 images per rank per exchange
    # (the host side of a NCCL call costs 0.2-0.6 ms at 2-8 ranks: one per step would make the loop host-bound)
    GATHER_EVERY = int(os.environ.get("YC_GATHER_EVERY", "16"))
    gather = DetectionGather(pipe.message(GATHER_ROWS).numel(), dev, every=GATHER_EVERY) if world > 1 else None
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = [torch.randn(args.bs, c, h, w, generator=g, device=dev).to(tdt) for c, (h, w) in zip(CH, SHAPES)]
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
            prev = pipe.submit(xs)
            if world > 1 and prev is not None:
                gather.gather_async(pipe.message(GATHER_ROWS, previous=True))
        else:
            pipe.run_device(xs)
            if world > 1:
                gather.gather_async(pipe.message(GATHER_ROWS), stream=pipe.tail_stream)

    def finish():
        if pipelined:
            pipe.drain()
            if world > 1:
                gather.gather_async(pipe.message(GATHER_ROWS))
        if world > 1:
            gather.flush()
            gather.wait()
        pipe.wait()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput: K steps between two events (all work of the K steps completes inside) ----------
    for _ in range(W):
        step()
    finish()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    t0.record()
    host_t0 = time.perf_counter()
    for i in range(K):
        step()
    host_loop = time.perf_counter() - host_t0
    finish()
    t1.record()
    sync_all()
    ms = t0.elapsed_time(t1)
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
    clocks = sampler.stop()
    ms = t0.elapsed_time(t1)
    head_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    tail_ms = statistics.mean(e[2].elapsed_time(e[3]) for e in ev)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
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
        if world > 1:
            gather.gather_async(pipe.message(GATHER_ROWS), stream=pipe.tail_stream)
    if overlap:
        out = pipe.drain_host()
    if world > 1:
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

    # ---- NMS latency at batch 1 (second metric of BASELINE.json): threshold/compaction + per-class NMS + letterbox
    # undo on one decoded 640x640 image (z resident in HBM), preallocated buffers, replayed as a CUDA graph ----------
    nms_lat = None
    if rank == 0:
        nms_lat = {}
        zsrc = PostBackbone(head, 8, SHAPES, tdt, INPUT_SHAPE, IMAGE_SHAPE, True, CONF, IOU, dev, use_graph=False,
                            fused=False)
        zsrc.run_device([x[:8].contiguous() for x in xs])
        for name, conf, iou in (("c2_conf0.25_iou0.45", CONF, IOU), ("c3_conf0.001_iou0.65", 0.001, 0.65)):
            p1 = PostBackbone(head, 1, SHAPES, tdt, INPUT_SHAPE, IMAGE_SHAPE, True, conf, iou, dev, use_graph=False,
                              fused=False)
            p1.z.copy_(zsrc.z[3:4])

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
            nms_lat[name] = {"p50_ms": statistics.median(lat), "detections": int(p1.meta[0].item())}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured"
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_launch = (BYTES_PER_IMG_FUSED if pipe.fused else BYTES_PER_IMG)[args.dtype] * args.bs
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
                "kernel_share_of_step": head_ms / (ms / K),
                "kernel_timing": f"CUDA events around each launch in a second pass of the same {K} steps issued call by call "
                                 f"({eager_ms / K:.4f} ms per step; the graph-replayed timed region above cannot hold "
                                 f"per-kernel events)",
                # the three NMS kernels of a step, timed on their own stream (they run next to the NEXT step's head
                # kernel when pipelined, which stretches them); kernel_share_serialised comes from a single-stream
                # pass of 20 steps and is what an ncu launch list (serialised) shows
                "nms_kernels_ms": tail_ms,
                "kernel_share_serialised": serial_share if serial_share else head_ms / (head_ms + tail_ms),
                "bytes_per_launch": bytes_launch,
                "flops_per_launch": FLOPS_PER_IMG * args.bs}
    if pipe.fused:
        # S3 moves 5.73 MB/img but still needs 1.462 GFLOP/img: the tensor pipe binds first
        tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        roofline.update({"bound": "tensor", "achieved": tflops, "peak": tpeak, "unit": "TFLOP/s", "frac": tflops / tpeak,
                         "hbm_gbs": achieved, "hbm_frac": achieved / hbm_peak})
    else:
        roofline.update({"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "tensor_tflops": tflops})
    line = {
        "metric": "post_backbone_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
        "config": dict(workload_config(args.bs, args.dtype), pipelining=pipelining_note(overlap),
                       **({"exchange": f"one NCCL all-gather per {GATHER_EVERY} steps: per rank and step the header + the first "
                                       f"{GATHER_ROWS} detection rows; every step's detections reach every rank inside the "
                                       f"timed region"} if world > 1 else {})),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": pipe.h2d_bytes(),
                "d2h_bytes_per_step": pipe.d2h_bytes(total_rows), "ms_per_step": e_ms / K},
        "gpu_launches": K * pipe.kernels_per_step,
        "roofline": roofline,
        "detections_per_step": n_det, "nms_latency_bs1": nms_lat,
    }
    if world == 1 and not args.no_cpu_baseline:
        ips, cores, sec, n = time_cpu_port(make_head(), 16, 5, 2, min_seconds=12.0)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"16 images of the same workload per pass, {n} timed passes over "
                                          f"{sec * n:.0f} s (median {sec:.3f} s per pass), torch CPU ops with {cores} threads"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
