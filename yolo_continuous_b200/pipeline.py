"""Post-backbone pipeline with preallocated buffers: neck feature maps -> detections.

This is the batched, allocation-free form of what reference detect.predict does after the
backbone (detect.py:227-234): head (conv + implicit + decode) -> threshold/compaction -> per-class
NMS -> letterbox undo.  Two entry points:

* run_device(features): inputs already on the device; results stay on the device.  The whole step
  (1 head kernel + the NMS kernels) can be replayed as one CUDA graph.  With overlap=True the NMS kernels of a
  batch run on a second stream (`tail_stream`) next to the head kernel of the following batch: the persistent
  head kernel leaves a quarter of each SM's registers and ~19 KB of its shared memory free for them, workspaces and
  outputs are double-buffered, and the results of a call are complete once `done_event` has fired.
* run_host(features_host): the user-facing call with HOST buffers -- pinned host feature maps are
  copied to the device, the step runs, and the per-image detection arrays are copied back.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class PostBackbone:
    def __init__(self, head, bs, shapes, dtype=torch.bfloat16, input_shape=(640, 640), image_shape=(640, 640),
                 letterbox_image=True, conf_thres=0.25, nms_thres=0.45, device="cuda:0", use_graph=True,
                 spec_rows=65536, fused=True, double_buffer=False, overlap=False, channels_last=False):
        self.head, self.bs, self.shapes = head, bs, [tuple(s) for s in shapes]
        self.device = torch.device(device)
        self.dtype = dtype
        self.nc, self.na, self.no = head.nc, head.na, head.no
        # channels_last: the feature maps are torch channels_last tensors ([bs, H, W, K] in memory, e.g. straight out of a
        # channels-last neck): the head GEMM reads them as a K-major operand (bf16 maps, tcgen05 path only)
        self.channels_last = bool(channels_last)
        if self.channels_last and dtype != torch.bfloat16:
            raise _lib.YcError("channels_last=True needs bfloat16 feature maps")
        # IDetect / IAuxDetect (no = nc + 5) and IBin (no = nc + 3 + 2 * (bin_count + 1): nets/ibin.py:20-21)
        self.ibin = hasattr(head, "w_bin_sigmoid")
        if not hasattr(head, "m") or head.no != head.nc + (3 + 2 * (head.bin_count + 1) if self.ibin else 5):
            raise _lib.YcError("PostBackbone drives IDetect / IAuxDetect / IBin heads")
        self.nl = len(self.shapes)
        dev = self.device
        self.rows = sum(self.na * h * w for h, w in self.shapes)
        self.ch = [head.m[i].weight.shape[1] for i in range(self.nl)]
        with torch.cuda.device(dev):
            fmt = torch.channels_last if self.channels_last else torch.contiguous_format
            self.x_dev = [torch.empty((bs, c, h, w), dtype=dtype, device=dev, memory_format=fmt)
                          for c, (h, w) in zip(self.ch, self.shapes)]
            self.no_out = self.nc + 5      # columns of a z row (IBin re-packs its 127 outputs to nc + 5, nets/ibin.py:72)
            self.z = torch.empty((bs, self.rows, self.no_out), dtype=torch.float32, device=dev)
            # Output message(s): [counts (bs) | offsets (bs+1) | pad] int32 header followed by the detection rows
            # [bs*rows, 7] fp32 in ONE allocation, so that a fixed-size prefix (header + the first `gather_rows`
            # rows) can be sent as a single all-gather / D2H copy.  Two of them when double-buffered (the
            # multi-GPU exchange of step i overlaps step i+1).
            self.hdr_ints = (2 * bs + 1 + 3) // 4 * 4
            self.overlap = bool(overlap)
            self.n_bufs = 2 if (double_buffer or overlap) else 1
            self.msgs = [torch.empty((self.hdr_ints * 4 + bs * self.rows * 28,), dtype=torch.uint8, device=dev)
                         for _ in range(self.n_bufs)]
            self.metas = [m[:self.hdr_ints * 4].view(torch.int32)[:2 * bs + 1] for m in self.msgs]
            self.rows_bufs = [m[self.hdr_ints * 4:].view(torch.float32).view(bs * self.rows, 7) for m in self.msgs]
            self.cur = 0
            n_ws = 2 if overlap else 1
            self.out_idxs = [torch.empty((bs * self.rows,), dtype=torch.int32, device=dev) for _ in range(n_ws)]
            self.wss = [torch.empty(_lib.lib.yc_nms_workspace_bytes(bs, self.rows, self.nc) + 1024, dtype=torch.uint8,
                                    device=dev) for _ in range(n_ws)]
            self.tail_stream = torch.cuda.Stream(device=dev) if overlap else None
            self.ev_head = [torch.cuda.Event() for _ in range(n_ws)]
            self.ev_tail = [torch.cuda.Event() for _ in range(n_ws)]
            hw = np.asarray(image_shape, dtype=np.int32).reshape(-1, 2)
            self.image_hw = torch.from_numpy(np.ascontiguousarray(hw)).to(dev)
            self.x_host = [torch.empty(t.shape, dtype=dtype, memory_format=fmt).pin_memory() for t in self.x_dev]
            self.spec_rows = min(spec_rows, bs * self.rows)
            self.meta_host = torch.empty((2 * bs + 1,), dtype=torch.int32).pin_memory()
            self.rows_host = torch.empty((bs * self.rows, 7), dtype=torch.float32).pin_memory() \
                if bs * self.rows <= (1 << 22) else torch.empty((1 << 22, 7), dtype=torch.float32).pin_memory()
        # descriptors (pointers filled per call for the head inputs)
        d = _lib.HeadDesc()
        d.kind, d.path = (_lib.YC_HEAD_IBIN if self.ibin else _lib.YC_HEAD_IDETECT), head.head_path
        if self.ibin:
            self._bins = head.w_bin_sigmoid.bins.to(device=dev, dtype=torch.float32).contiguous()
            d.bin_count, d.bins = head.bin_count, self._bins.data_ptr()
        d.x_dtype = _lib.YC_BF16 if dtype == torch.bfloat16 else _lib.YC_F32
        d.nl, d.na, d.no, d.bs = self.nl, self.na, self.no, bs
        d.x_channels_last = 1 if self.channels_last else 0
        self._blobs = []
        for i, (h, w) in enumerate(self.shapes):
            blob = head._blob((id(head.m[i]),), head.m[i], head.ia[i], head.im[i], dev)
            self._blobs.append(blob)
            lv = d.level[i]
            lv.blob, lv.K, lv.H, lv.W = blob.data_ptr(), self.ch[i], h, w
            lv.stride = float(head.stride[i])
            for j, v in enumerate(head.anchor_grid[i].reshape(-1).tolist()):
                lv.anchor_wh[j] = v
        d.z = self.z.data_ptr()
        self.desc = d
        self._weight_key = self._weights_version()
        p = _lib.NmsParams()
        p.bs, p.rows, p.row_stride, p.nc = bs, self.rows, self.no_out, self.nc
        p.conf_thres, p.nms_thres = float(conf_thres), float(nms_thres)
        p.write_corners, p.correct_boxes, p.letterbox = 0, 1, 1 if letterbox_image else 0
        p.input_h, p.input_w = int(input_shape[0]), int(input_shape[1])
        p.image_hw, p.image_hw_stride = self.image_hw.data_ptr(), (2 if hw.shape[0] > 1 else 0)
        p.box_div_w, p.box_div_h = float(input_shape[1]), float(input_shape[0])
        self.nms_params = p
        # fused mode: one head kernel whose epilogue emits the NMS candidates (z never written) + 3 NMS kernels
        # (bucket, per-segment NMS, finish); otherwise head kernel (writes z) + threshold/compaction + the same 3.
        # (memsets are not kernels)
        self.fused = fused and dtype == torch.bfloat16
        if self.overlap and not self.fused:
            raise _lib.YcError("overlap=True needs the fused step (bf16 feature maps)")
        self.kernels_per_step = 4 if self.fused else 5
        self._graphs = {}
        self._pgraphs = {}
        self._in_flight = False
        self._n_sub = 0
        self.exchange, self.xchg_stream = None, None
        self._eager_dirty = False
        self._hp = None
        self.use_graph = use_graph and not self.overlap
        if self.overlap:
            with torch.cuda.device(dev):
                for i in range(2):
                    _lib.check(_lib.lib.yc_nms_workspace_reset(C.byref(p), self.wss[i].data_ptr(), self.wss[i].numel(),
                                                               C.c_void_p(self.tail_stream.cuda_stream)),
                               "yc_nms_workspace_reset")
                    self.ev_tail[i].record(self.tail_stream)

    # ---- head parameters: the packed blobs follow the parameters' version counters --------------------------------
    def _weights_version(self):
        h = self.head
        return tuple(p._version for i in range(self.nl)
                     for p in (h.m[i].weight, h.m[i].bias, h.ia[i].implicit, h.im[i].implicit) if p is not None)

    def refresh_weights(self, force=False):
        """Re-pack the head parameters if they changed since the descriptor was built (in-place updates, e.g.
        load_state_dict, bump the version counters checked here on every call; call with force=True after replacing
        parameter tensors).  Captured graphs are dropped: they hold the old blob pointers."""
        key = self._weights_version()
        if not force and key == self._weight_key:
            return False
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)    # work in flight still reads the old blobs
            if force:
                self.head._packed.clear()
            self._blobs = []
            for i in range(self.nl):
                blob = self.head._blob((id(self.head.m[i]),), self.head.m[i], self.head.ia[i], self.head.im[i], self.device)
                self._blobs.append(blob)
                self.desc.level[i].blob = blob.data_ptr()
        self._graphs.clear()
        self._pgraphs.clear()
        self._weight_key = key
        return True

    @property
    def ws(self):
        return self.wss[self.cur if self.overlap else 0]

    @property
    def out_idx(self):
        return self.out_idxs[self.cur if self.overlap else 0]

    @property
    def done_event(self):
        """Fires when the results of the latest run_device call are complete (overlap mode)."""
        return self.ev_tail[self.cur]

    def wait(self):
        """Make the current stream wait for the results of the latest run_device call (no-op without overlap)."""
        if self.overlap:
            torch.cuda.current_stream(self.device).wait_event(self.ev_tail[self.cur])

    @property
    def meta(self):
        return self.metas[self.cur]

    @property
    def out_rows(self):
        return self.rows_bufs[self.cur]

    def message(self, gather_rows, previous=False):
        """Fixed-size prefix of the current output buffer: header + the first `gather_rows` detection rows
        (previous=True: of the batch before, i.e. the one whose results submit() just returned)."""
        return self.msgs[self.cur ^ 1 if previous else self.cur][:self.hdr_ints * 4 + gather_rows * 28]

    # ---- device path -----------------------------------------------------------------------
    def _check(self, i, x):
        ok = x.is_contiguous(memory_format=torch.channels_last) if self.channels_last else x.is_contiguous()
        if x.dtype != self.dtype or tuple(x.shape) != tuple(self.x_dev[i].shape) or not ok:
            raise _lib.YcError(f"level {i}: expected {'channels-last' if self.channels_last else 'contiguous'} "
                               f"{tuple(self.x_dev[i].shape)} {self.dtype}")

    def _launch(self, features, head_events=None):
        for i, x in enumerate(features):
            self._check(i, x)
            self.desc.level[i].x = x.data_ptr()
        s = _lib.stream_ptr(self.device)
        m = self.meta.data_ptr()
        if self.fused and (self.overlap or head_events is not None):
            # head and NMS halves as two calls: the NMS half on the tail stream (overlap) and/or the head half
            # bracketed by the caller's events (bench.py's per-kernel timing)
            main = torch.cuda.current_stream(self.device)
            c = self.cur
            if self.overlap:
                main.wait_event(self.ev_tail[c])    # the tail that used this workspace/output two calls ago
            if head_events is not None:
                head_events[0].record(main)
            # overlap: the workspace counters were cleared on the tail stream (at construction / behind the NMS
            # kernels that used them last), so the head stream carries nothing but the head kernel
            head_fn = _lib.lib.yc_detect_fused_head_noreset if self.overlap else _lib.lib.yc_detect_fused_head
            _lib.check(head_fn(C.byref(self.desc), C.byref(self.nms_params), self.ws.data_ptr(), self.ws.numel(), s),
                       "yc_detect_fused_head")
            if head_events is not None:
                head_events[1].record(main)
            tail = main
            if self.overlap:
                self.ev_head[c].record(main)
                tail = self.tail_stream
                tail.wait_event(self.ev_head[c])
            if head_events is not None and len(head_events) > 2:
                head_events[2].record(tail)
            _lib.check(_lib.lib.yc_nms_from_candidates(C.byref(self.nms_params), self.ws.data_ptr(), self.ws.numel(),
                                                       self.out_rows.data_ptr(), self.out_idx.data_ptr(), m,
                                                       m + 4 * self.bs, C.c_void_p(tail.cuda_stream)),
                       "yc_nms_from_candidates")
            if head_events is not None and len(head_events) > 2:
                head_events[3].record(tail)
            if self.overlap:
                _lib.check(_lib.lib.yc_nms_workspace_reset(C.byref(self.nms_params), self.ws.data_ptr(), self.ws.numel(),
                                                           C.c_void_p(tail.cuda_stream)), "yc_nms_workspace_reset")
                if self.exchange is not None:
                    import os
                    dbg = int(os.environ.get("YC_XCHG_DEBUG", "0"))
                    tev = getattr(self, "_xchg_events", None)     # timing experiments (tools/xchg_time.py)
                    if tev is not None:
                        tev.append([torch.cuda.Event(enable_timing=True) for _ in range(3)])
                        tev[-1][0].record(tail)
                    if dbg < 2:
                        self.exchange.push(self.msgs[c], tail)
                    if tev is not None:
                        tev[-1][1].record(tail)
                    if dbg < 1:
                        self.exchange.wait(tail, lag=1)
                    if tev is not None:
                        tev[-1][2].record(tail)
                self.ev_tail[c].record(tail)
                self._eager_dirty = True
            return
        if self.fused:
            rc = _lib.lib.yc_detect_fused(C.byref(self.desc), C.byref(self.nms_params), self.ws.data_ptr(),
                                          self.ws.numel(), self.out_rows.data_ptr(), self.out_idx.data_ptr(),
                                          m, m + 4 * self.bs, s)
            if rc != _lib.YC_ERR_UNSUPPORTED:
                _lib.check(rc, "yc_detect_fused")
                return
            self.fused, self.kernels_per_step = False, 5   # shape does not fit the tcgen05 kernel: two-call path
        if head_events is not None:
            head_events[0].record()
        _lib.check(_lib.lib.yc_head_forward(C.byref(self.desc), s), "yc_head_forward")
        if head_events is not None:
            head_events[1].record()
            if len(head_events) > 2:
                head_events[2].record()
        _lib.check(_lib.lib.yc_nms_batched(self.z.data_ptr(), C.byref(self.nms_params), self.ws.data_ptr(),
                                           self.ws.numel(), self.out_rows.data_ptr(), self.out_idx.data_ptr(),
                                           m, m + 4 * self.bs, s), "yc_nms_batched")
        if head_events is not None and len(head_events) > 2:
            head_events[3].record()

    def run_device(self, features, head_events=None):
        """features: list of [bs, ch_i, H_i, W_i] device tensors.  Returns device views
        (rows [bs*rows,7] capacity, idx, counts [bs], offsets [bs+1]); with overlap=True they are complete once
        `done_event` has fired (`wait()` orders the current stream behind it)."""
        self.refresh_weights()
        with torch.cuda.device(self.device):
            if self._in_flight:
                raise _lib.YcError("run_device() while a submit() batch is in flight: call drain() first")
            if self.n_bufs > 1:
                self.cur ^= 1
            ptrs = tuple(x.data_ptr() for x in features) + (self.cur,)
            if not self.use_graph or head_events is not None:
                self._launch(features, head_events)
            else:
                g = self._graphs.get(ptrs)
                if g is None:
                    self._launch(features)          # warm-up outside capture (lazy module/attribute setup)
                    torch.cuda.current_stream().synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may call CUDA
                        self._launch(features)
                    if len(self._graphs) > 8:
                        self._graphs.clear()
                    self._graphs[ptrs] = g
                g.replay()
        bs = self.bs
        return self.out_rows, self.out_idx, self.meta[:bs], self.meta[bs:]

    # ---- software-pipelined device path: one CUDA graph per step ------------------------------------------------
    def submit(self, features):
        """Pipelined form of run_device for a stream of batches (needs overlap=True): ONE graph launch runs the head
        kernel of THIS batch and, next to it on the tail stream, the NMS kernels of the PREVIOUS batch (fork / join
        inside the graph), so the host pays one launch per step instead of ~10 calls (measured: 70 us of enqueue
        work per step for the eager two-stream path against a 89 us step).  Returns the device views of the previous
        batch's results -- complete for work enqueued after this call on the current stream -- or None on the first
        call; drain() returns the last batch's.  `features` must alternate between at most a few fixed sets of
        tensors (graphs are keyed by the input pointers)."""
        if not (self.overlap and self.fused):
            raise _lib.YcError("submit() needs overlap=True and the fused step")
        self.refresh_weights()
        with torch.cuda.device(self.device):
            if self._eager_dirty:   # order behind NMS kernels an earlier run_device() left on the tail stream
                torch.cuda.current_stream().wait_event(self.ev_tail[0])
                torch.cuda.current_stream().wait_event(self.ev_tail[1])
                self._eager_dirty = False
            prev = self.cur if self._in_flight else None
            self.cur ^= 1
            c = self.cur
            # The first step of a stream of batches has no previous batch: the head kernel alone is launched (eagerly, one
            # call), so that results an earlier eager run_device() call left in the other output buffer stay untouched
            if not self._in_flight:
                for i, x in enumerate(features):
                    self._check(i, x)
                self._pipelined_step(features, c, with_tail=False)
                self._in_flight = True
                self._n_sub = 1
                return None
            self._n_sub += 1
            # multi-GPU: from the third batch on the graph has a third branch that pushes the results of the batch before
            # the previous one (complete since the previous graph) to all ranks and awaits earlier messages
            with_push = self.exchange is not None and self._n_sub >= 3
            ptrs = tuple(x.data_ptr() for x in features)
            g = self._pgraphs.get(ptrs + (c, with_push))
            if g is None:
                # a new set of input tensors: capture every variant of the step for it at once (both output buffers, with
                # and without the exchange branch), so that no later step of the stream has to stop for a capture
                for i, x in enumerate(features):
                    self._check(i, x)
                torch.cuda.current_stream().wait_event(self.ev_tail[0])
                torch.cuda.current_stream().wait_event(self.ev_tail[1])
                torch.cuda.current_stream().synchronize()
                if len(self._pgraphs) > 32:
                    self._pgraphs.clear()
                for cc in (0, 1):
                    for wp in ((False, True) if self.exchange is not None else (False,)):
                        gg = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gg, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may call CUDA
                            self._pipelined_step(features, cc, with_push=wp)
                        self._pgraphs[ptrs + (cc, wp)] = gg
                g = self._pgraphs[ptrs + (c, with_push)]
            g.replay()
            self._in_flight = True
        return None if prev is None else self._views(prev)

    def drain(self):
        """NMS kernels of the last submitted batch (eager); returns its result views, complete on the current stream.
        With an exchange attached the batches not pushed yet (the last two) are pushed, in order; follow with
        `exchange.wait_all()` to have every rank's messages."""
        if not self._in_flight:
            return None
        with torch.cuda.device(self.device):
            c = self.cur
            s = torch.cuda.current_stream()
            if self.exchange is not None and self._n_sub >= 2:
                # the batch before the last is complete already: its push runs next to the last batch's NMS kernels
                self.xchg_stream.wait_stream(s)
                self.exchange.push(self.msgs[1 - c], self.xchg_stream)
            self._tail(c, s)
            if self.exchange is not None:
                if self._n_sub >= 2:
                    s.wait_stream(self.xchg_stream)
                self.exchange.push(self.msgs[c], s)
            self._in_flight = False
        return self._views(c)

    def _views(self, c):
        bs = self.bs
        return self.rows_bufs[c], self.out_idxs[c], self.metas[c][:bs], self.metas[c][bs:]

    def _tail(self, c, stream):
        m = self.metas[c].data_ptr()
        sp = C.c_void_p(stream.cuda_stream)
        _lib.check(_lib.lib.yc_nms_from_candidates(C.byref(self.nms_params), self.wss[c].data_ptr(), self.wss[c].numel(),
                                                   self.rows_bufs[c].data_ptr(), self.out_idxs[c].data_ptr(), m,
                                                   m + 4 * self.bs, sp), "yc_nms_from_candidates")
        _lib.check(_lib.lib.yc_nms_workspace_reset(C.byref(self.nms_params), self.wss[c].data_ptr(), self.wss[c].numel(), sp),
                   "yc_nms_workspace_reset")

    def attach_exchange(self, exchange):
        """Multi-GPU: `exchange` (parallel.PeerExchange built with this pipeline's hdr_ints / bs) receives every batch's
        detections.  Pipelined steps (submit / drain): the push and wait kernels form a third branch of the per-step CUDA
        graph, next to the head kernel and the NMS kernels (behind the NMS kernels they would lengthen the critical
        branch: each small kernel costs 5-10 us next to the persistent head kernel): the graph of batch i pushes the
        results of batch i - 2 and awaits earlier messages; drain() pushes the last two batches.  The i-th batch of a
        stream carries the i-th sequence number; call exchange.wait_all() after drain().  Eager run_device() calls push and
        await on the tail stream behind the NMS kernels."""
        if exchange is not None and (exchange.hdr_ints != self.hdr_ints or exchange.bs != self.bs):
            raise _lib.YcError("exchange was built for another message layout")
        if self._in_flight:
            raise _lib.YcError("attach_exchange() while a submit() batch is in flight: call drain() first")
        self.exchange = exchange
        self._pgraphs.clear()
        if exchange is not None and self.xchg_stream is None:
            self.xchg_stream = torch.cuda.Stream(device=self.device)

    def _pipelined_step(self, features, c, with_tail=True, with_push=False):
        """(captured) head of the current batch into workspace c  ||  NMS kernels of the previous batch (workspace
        1-c)  ||  (with_push) exchange of the batch before that (output buffer c); with_tail=False (first step of a
        stream of batches): the head kernel only."""
        main = torch.cuda.current_stream()
        side = self.tail_stream
        if with_tail:
            side.wait_stream(main)                               # fork
        if with_push:
            self.xchg_stream.wait_stream(main)
        for i, x in enumerate(features):
            self.desc.level[i].x = x.data_ptr()
        _lib.check(_lib.lib.yc_detect_fused_head_noreset(C.byref(self.desc), C.byref(self.nms_params),
                                                         self.wss[c].data_ptr(), self.wss[c].numel(),
                                                         C.c_void_p(main.cuda_stream)), "yc_detect_fused_head")
        if with_tail:
            self._tail(1 - c, side)
        if with_push:
            self.exchange.push(self.msgs[c], self.xchg_stream)
            self.exchange.wait(self.xchg_stream, lag=1)
            main.wait_stream(self.xchg_stream)
        if with_tail:
            main.wait_stream(side)                               # join

    # ---- host path (the e2e call) -----------------------------------------------------------------
    def run_host(self, features_host=None):
        """features_host: list of pinned host tensors (default: self.x_host).  Returns the reference's
        result type: list with None or ndarray[n,7] (y1,x1,y2,x2 px, obj, class_conf, class_id) per image."""
        src = self.x_host if features_host is None else features_host
        with torch.cuda.device(self.device):
            for d_, h_ in zip(self.x_dev, src):
                d_.copy_(h_, non_blocking=True)
            rows, _, _, _ = self.run_device(self.x_dev)
            self.wait()
            self.meta_host.copy_(self.meta, non_blocking=True)
            spec = self.spec_rows
            self.rows_host[:spec].copy_(rows[:spec], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            off = self.meta_host[self.bs:].numpy()
            total = int(off[-1])
            if total > spec:
                if total > self.rows_host.shape[0]:
                    raise _lib.YcError("detections exceed the pinned host buffer")
                self.rows_host[spec:total].copy_(rows[spec:total], non_blocking=True)
                torch.cuda.current_stream().synchronize()
        host = self.rows_host.numpy()
        return [None if off[b + 1] == off[b] else host[off[b]:off[b + 1]].copy() for b in range(self.bs)]

    # ---- host path, software pipelined: H2D of batch i+1 under the kernels and the D2H of batch i -----------------
    def submit_host(self, features_host=None):
        """Pipelined run_host for a stream of batches (needs overlap=True): the pinned host maps of THIS batch are
        copied on a copy stream into the second set of device buffers while the kernels and the result read-back of
        the previous batch are still in flight, so the PCIe link never idles between batches.  Returns the previous
        batch's detections (reference result type, list of None | ndarray[n,7]) or None on the first call;
        drain_host() returns the last batch's."""
        if not self.overlap:
            raise _lib.YcError("submit_host() needs overlap=True")
        src = self.x_host if features_host is None else features_host
        with torch.cuda.device(self.device):
            if self._hp is None:
                dev = self.device
                self._hp = {
                    "x": [self.x_dev, [torch.empty_like(t) for t in self.x_dev]],
                    "copy": torch.cuda.Stream(device=dev),
                    "h2d": [torch.cuda.Event(), torch.cuda.Event()], "xfree": [torch.cuda.Event(), torch.cuda.Event()],
                    "d2h": [torch.cuda.Event(), torch.cuda.Event()],
                    "meta": [self.meta_host, torch.empty_like(self.meta_host).pin_memory()],
                    "rows": [self.rows_host[:self.spec_rows], torch.empty((self.spec_rows, 7), dtype=torch.float32).pin_memory()],
                    "k": 0, "pending": None}
            hp = self._hp
            k = hp["k"] = hp["k"] ^ 1
            main = torch.cuda.current_stream()
            with torch.cuda.stream(hp["copy"]):
                hp["copy"].wait_event(hp["xfree"][k])           # the head kernel that last read this set is done
                for d_, h_ in zip(hp["x"][k], src):
                    d_.copy_(h_, non_blocking=True)
                hp["h2d"][k].record(hp["copy"])
            main.wait_event(hp["h2d"][k])
            rows, _, _, _ = self.run_device(hp["x"][k])          # head on `main`, NMS kernels on the tail stream
            hp["xfree"][k].record(main)
            c = self.cur
            with torch.cuda.stream(self.tail_stream):            # read-back in order behind the NMS kernels
                hp["meta"][k].copy_(self.metas[c], non_blocking=True)
                hp["rows"][k].copy_(rows[:self.spec_rows], non_blocking=True)
                hp["d2h"][k].record(self.tail_stream)
            prev, hp["pending"] = hp["pending"], (k, c)
        return None if prev is None else self._collect_host(*prev)

    def drain_host(self):
        if self._hp is None or self._hp["pending"] is None:
            return None
        prev, self._hp["pending"] = self._hp["pending"], None
        return self._collect_host(*prev)

    def _collect_host(self, k, c):
        hp = self._hp
        hp["d2h"][k].synchronize()
        off = hp["meta"][k][self.bs:].numpy()
        total = int(off[-1])
        host = hp["rows"][k].numpy()
        if total > self.spec_rows:   # rare: more detections than the speculative read-back holds
            extra = self.rows_bufs[c][:total].cpu().numpy()
            return [None if off[b + 1] == off[b] else extra[off[b]:off[b + 1]].copy() for b in range(self.bs)]
        return [None if off[b + 1] == off[b] else host[off[b]:off[b + 1]].copy() for b in range(self.bs)]

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.x_host)

    def d2h_bytes(self, total_rows):
        return self.meta_host.numel() * 4 + max(self.spec_rows, total_rows) * 28
