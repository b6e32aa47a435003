"""Plain Detect head (reference nets/detect.py): three 1x1 convs, raw NCHW maps, P5 first.

This is the head the shipped YAMLs use (Variant A).  `forward` keeps the reference
contract (raw conv maps, torch convolutions).  `forward_decoded` is the B200 path for inference: the same
fused kernel as IDetect (1x1 conv as a tcgen05 GEMM: bf16 maps directly, fp32 maps through the fp16 hi/lo split, sigmoid and
box decode in the epilogue) with Variant A's normalised-box decode -- it returns what
`decode_box(self(x), anchors, anchors_mask, num_classes, image_size)` returns (detect.py:29-87) without ever
materialising the raw maps.
"""
import ctypes as C

import numpy as np
import torch
from torch import nn

from .. import _lib
from ._head import HeadBase


class Detect(nn.Module):
    head_path = _lib.YC_PATH_AUTO

    def __init__(self, num_classes=80, anchors=(), ch=()):
        super().__init__()
        self.num_classes = num_classes
        self.len_output = num_classes + 5
        self.num_layers = len(anchors)
        self.num_anchors_each_layer = len(anchors[0]) // 2
        self.na = self.num_anchors_each_layer   # read by the shared parameter packing (HeadBase._blob)
        n = self.num_anchors_each_layer * self.len_output
        self.yolo_head_P3 = nn.Conv2d(ch[0], n, 1)
        self.yolo_head_P4 = nn.Conv2d(ch[1], n, 1)
        self.yolo_head_P5 = nn.Conv2d(ch[2], n, 1)
        for mod in self.modules():  # reference nets/detect.py:18-25
            if isinstance(mod, (nn.Conv2d, nn.Linear)):
                nn.init.normal_(mod.weight, 0, 0.01)
        self._packed = {}

    def forward(self, x):
        return [self.yolo_head_P5(x[2]), self.yolo_head_P4(x[1]), self.yolo_head_P3(x[0])]

    _blob = HeadBase._blob   # parameter packing cached on the parameters' version counters

    def forward_decoded(self, x, anchors, anchors_mask, image_size=(640, 640)):
        """x: [P3, P4, P5] neck maps (CUDA, float32 or bfloat16).  Returns the list decode_box would return for
        forward(x): per level (P5, P4, P3) a [bs, na*H*W, 5+nc] float32 tensor with normalised xywh and
        sigmoid scores; the three are consecutive views of one buffer, so torch.cat(outputs, 1) is free
        (`outputs[0]._base` is the [bs, sum, 5+nc] tensor the reference concatenates, detect.py:232)."""
        convs = [self.yolo_head_P5, self.yolo_head_P4, self.yolo_head_P3]
        xs = [x[2], x[1], x[0]]
        anchors = np.asarray(anchors, dtype=np.float64).reshape(-1, 2)
        na, no = self.num_anchors_each_layer, self.len_output
        x0 = xs[0]
        for t in xs:
            _lib.require_cuda(t, "Detect input")
            if t.dim() != 4 or t.dtype != x0.dtype or t.device != x0.device or t.shape[0] != x0.shape[0]:
                raise _lib.YcError("Detect inputs must be 4-D NCHW tensors sharing dtype, device and batch size")
        if x0.dtype not in (torch.float32, torch.bfloat16):
            raise _lib.YcError(f"unsupported feature-map dtype {x0.dtype}: use float32 or bfloat16")
        dev, bs = x0.device, x0.shape[0]
        d = _lib.HeadDesc()
        d.kind, d.path = _lib.YC_HEAD_IDETECT, self.head_path
        d.x_dtype = _lib.YC_F32 if x0.dtype == torch.float32 else _lib.YC_BF16
        d.nl, d.na, d.no, d.bs = 3, na, no, bs
        keep, rows = [], []
        with torch.cuda.device(dev):
            for i in range(3):
                t = xs[i].contiguous()
                keep.append(t)
                _, k, h, w = t.shape
                if k != convs[i].weight.shape[1]:
                    raise _lib.YcError(f"level {i}: expected {convs[i].weight.shape[1]} channels, got {k}")
                if len(anchors_mask[i]) != na:
                    raise _lib.YcError(f"level {i}: anchors_mask has {len(anchors_mask[i])} anchors, the head has {na}")
                lv = d.level[i]
                lv.x, lv.blob = t.data_ptr(), self._blob((id(convs[i]),), convs[i], None, None, dev).data_ptr()
                lv.K, lv.H, lv.W = k, h, w
                # xy = (2s - 0.5 + grid) / (W, H);  wh = (2s)^2 * (anchor / stride) / (W, H) with
                # stride_w = image_size[0] / W and stride_h = image_size[0] / H (detect.py:38-39,81-84)
                lv.stride, lv.stride_y = 1.0 / w, 1.0 / h
                for j, (aw, ah) in enumerate(anchors[anchors_mask[i]]):
                    lv.anchor_wh[2 * j], lv.anchor_wh[2 * j + 1] = aw / image_size[0], ah / image_size[0]
                rows.append(na * h * w)
            z = torch.empty((bs, sum(rows), no), dtype=torch.float32, device=dev)
            d.z = z.data_ptr()
            _lib.check(_lib.lib.yc_head_forward(C.byref(d), _lib.stream_ptr(dev)), "yc_head_forward")
        outs, r0 = [], 0
        for r in rows:
            outs.append(z[:, r0:r0 + r])
            r0 += r
        return outs
