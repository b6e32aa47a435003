"""Shared host logic of the B200 detection heads.

One call of `yc_head_forward` (include/yc_b200.h) performs, for every pyramid level at once,
what the reference spreads over ~16 torch kernels per level:
ImplicitA add -> 1x1 conv -> ImplicitM mul -> view/permute -> sigmoid -> xy/wh decode -> cat.
"""
import ctypes as C

import torch
from torch import nn

from .. import _lib
from .common import ImplicitA, ImplicitM


class HeadBase(nn.Module):
    """State and plumbing common to IDetect / IAuxDetect / IBin.

    Attribute names (`nc no nl na grid stride export anchors anchor_grid m ia im`) and the
    state_dict layout follow the reference (nets/idetect.py:11-24) so its checkpoints load.
    """
    stride = None   # set by the model builder, as in the reference (nets/idetect.py:8)
    export = False  # nets/idetect.py:9
    head_path = _lib.YC_PATH_AUTO   # which kernel family runs the conv (see include/yc_b200.h)
    return_raw = True               # eval forward returns (z, raw list) as the reference does
    # float32 feature maps keep float32 grade ("exact", 1e-5 parity): the tensor-core kernel that splits both operands into
    # fp16 hi/lo parts (head_tcs_kernel), or the FFMA kernel on the generic path.  "bf16" casts them to bfloat16 first and
    # takes the bf16 tcgen05 kernels (1e-3 parity, the precision class of the TF32 convolutions torch uses on a GPU by
    # default, ~3x faster); bfloat16 maps always take the bf16 tcgen05 kernels.
    fp32_maps = "exact"

    def _init_common(self, nc, anchors, no):
        self.nc = nc
        self.no = no
        self.nl = len(anchors)
        self.na = len(anchors[0]) // 2
        if self.nl > _lib.YC_MAX_LEVELS or self.na > _lib.YC_MAX_ANCHORS:
            raise ValueError(f"at most {_lib.YC_MAX_LEVELS} levels x {_lib.YC_MAX_ANCHORS} anchors are supported")
        self.grid = [torch.zeros(1)] * self.nl
        a = torch.tensor(anchors).float().view(self.nl, -1, 2)
        self.register_buffer('anchors', a)
        self.register_buffer('anchor_grid', a.clone().view(self.nl, 1, -1, 1, 1, 2))
        self._packed = {}

    def _make_lead(self, ch):
        n = self.no * self.na
        self.m = nn.ModuleList(nn.Conv2d(c, n, 1) for c in ch)
        self.ia = nn.ModuleList(ImplicitA(c) for c in ch)
        self.im = nn.ModuleList(ImplicitM(n) for _ in ch)

    @staticmethod
    def _make_grid(nx=20, ny=20):
        yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
        return torch.stack((xv, yv), 2).view((1, 1, ny, nx, 2)).float()

    # ---- parameter packing, cached on the parameters' version counters -------------------
    def _blob(self, tag, conv, ia, im, device):
        parts = [conv.weight, conv.bias] + ([ia.implicit] if ia is not None else []) + \
                ([im.implicit] if im is not None else [])
        key = tuple((p.data_ptr(), p._version) for p in parts if p is not None) + (str(device),)
        hit = self._packed.get(tag)
        if hit is not None and hit[0] == key:
            return hit[1]
        n, k = conv.weight.shape[0], conv.weight.shape[1]
        if conv.kernel_size != (1, 1) or conv.groups != 1:
            raise _lib.YcError("head convolutions must be 1x1, groups=1")
        blob = torch.empty(_lib.lib.yc_head_pack_bytes(n, k), dtype=torch.uint8, device=device)

        def dev32(t):
            return None if t is None else t.detach().to(device=device, dtype=torch.float32).contiguous()

        w, b = dev32(conv.weight), dev32(conv.bias)
        a_, m_ = dev32(ia.implicit if ia is not None else None), dev32(im.implicit if im is not None else None)
        _lib.check(_lib.lib.yc_head_pack(w.data_ptr(), b.data_ptr() if b is not None else None,
                                         a_.data_ptr() if a_ is not None else None,
                                         m_.data_ptr() if m_ is not None else None,
                                         n, k, self.na, blob.data_ptr(), _lib.stream_ptr(device)), "yc_head_pack")
        self._packed[tag] = (key, blob)
        return blob

    # ---- one C-ABI call for all levels --------------------------------------------------------
    def _run(self, xs, convs, ias, ims, kind, want_z, want_raw, no_out=None, bins=None, bin_count=0):
        nl = len(xs)
        if self.fp32_maps == "bf16" and not (torch.is_grad_enabled() and any(t.requires_grad for t in xs)):
            xs = [t.to(torch.bfloat16) if isinstance(t, torch.Tensor) and t.dtype == torch.float32 else t for t in xs]
        elif self.fp32_maps not in ("exact", "bf16"):
            raise ValueError(f"fp32_maps must be 'exact' or 'bf16', got {self.fp32_maps!r}")
        x0 = xs[0]
        for t in xs:
            if not isinstance(t, torch.Tensor) or t.dim() != 4:
                raise _lib.YcError("head inputs must be 4-D NCHW tensors")
            _lib.require_cuda(t, "head input")
            if t.dtype != x0.dtype or t.device != x0.device or t.shape[0] != x0.shape[0]:
                raise _lib.YcError("head inputs must share dtype, device and batch size")
            if t.requires_grad and torch.is_grad_enabled():
                raise NotImplementedError("the B200 head path is inference-only (no backward kernels)")
        if x0.dtype == torch.float32:
            xdt = _lib.YC_F32
        elif x0.dtype == torch.bfloat16:
            xdt = _lib.YC_BF16
        else:
            raise _lib.YcError(f"unsupported feature-map dtype {x0.dtype}: use float32 or bfloat16")
        dev, bs = x0.device, x0.shape[0]
        # channels-last bf16 maps (e.g. the output of a channels-last RepConv) are consumed as they are: the tcgen05
        # kernels read them as a K-major operand; anything else is taken as NCHW
        nhwc = xdt == _lib.YC_BF16 and self.head_path != _lib.YC_PATH_GENERIC and all(
            not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last) and t.shape[1] % 8 == 0 for t in xs)
        d = _lib.HeadDesc()
        d.kind, d.path, d.x_dtype = kind, self.head_path, xdt
        d.x_channels_last = 1 if nhwc else 0
        d.nl, d.na, d.no, d.bin_count, d.bs = nl, self.na, self.no, bin_count, bs
        keep = []
        raws, rows = [], 0
        with torch.cuda.device(dev):
            for i in range(nl):
                x = xs[i] if nhwc else xs[i].contiguous()
                keep.append(x)
                _, k, h, w = x.shape
                if k != convs[i].weight.shape[1]:
                    raise _lib.YcError(f"level {i}: expected {convs[i].weight.shape[1]} channels, got {k}")
                blob = self._blob((id(convs[i]),), convs[i], ias[i] if ias else None, ims[i] if ims else None, dev)
                lv = d.level[i]
                lv.x, lv.blob = x.data_ptr(), blob.data_ptr()
                lv.K, lv.H, lv.W = k, h, w
                if want_z:
                    lv.stride = self._stride_list()[i]  # TypeError when stride was never set, as the reference
                for j, v in enumerate(self._anchor_list(i)):
                    lv.anchor_wh[j] = v
                if want_raw:
                    r = torch.empty((bs, self.na, h, w, self.no), dtype=torch.float32, device=dev)
                    raws.append(r)
                    lv.raw = r.data_ptr()
                rows += self.na * h * w
            z = None
            if want_z:
                z = torch.empty((bs, rows, no_out or self.no), dtype=torch.float32, device=dev)
                d.z = z.data_ptr()
            if bins is not None:
                bins = bins.to(device=dev, dtype=torch.float32).contiguous()
                d.bins = bins.data_ptr()
            _lib.check(_lib.lib.yc_head_forward(C.byref(d), _lib.stream_ptr(dev)), "yc_head_forward")
        return z, raws

    def _stride_list(self):
        st = self.stride
        if not isinstance(st, torch.Tensor):
            return [float(st[i]) for i in range(self.nl)]   # None: TypeError, as `self.stride[i]` in the reference
        key = (st.data_ptr(), st._version)
        hit = self._packed.get("stride")
        if hit is None or hit[0] != key:
            hit = (key, [float(v) for v in st.reshape(-1).tolist()])
            self._packed["stride"] = hit
        return hit[1]

    def _anchor_list(self, i):
        # anchor_grid lives on the device: reading it back is a stream synchronisation, so the host copy is cached on the
        # buffer's identity and version (a reloaded or rescaled anchor_grid is read again)
        ag = self.anchor_grid
        key = (ag.data_ptr(), ag._version)
        hit = self._packed.get("anchors")
        if hit is None or hit[0] != key:
            hit = (key, ag.reshape(self.nl, -1).tolist())
            self._packed["anchors"] = hit
        return hit[1][i]

    def _update_grid_cache(self, i, ny, nx, device):
        # attribute kept for compatibility with code that inspects `head.grid` (nets/idetect.py:37-38);
        # the kernels derive the cell index from the row number and never read it
        if self.grid[i].shape[2:4] != (ny, nx):
            self.grid[i] = self._make_grid(nx, ny).to(device)
