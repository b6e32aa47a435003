"""Letterbox preprocessing on the device (drop-in for reference image_enhance/letter_box.py and the
prepare_test_image step of detect.py:16-26).

`letterbox_batch` is the B200 path: a batch of decoded uint8 HWC images (any sizes) becomes the network input
[bs, 3, H, W] (float32 or bfloat16, /255, 114-padded, channel order kept) in one kernel, bit-identical to
cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder + numpy.  `LetterBox` keeps the reference class signature for single
images (uint8 HWC in, uint8 HWC out, labels shifted) on top of the same kernel.
"""
import ctypes as C
from random import Random

import numpy as np
import torch

from .. import _lib


class _LbImage(C.Structure):   # yc_letterbox_image of include/yc_b200.h
    _fields_ = [("src", C.c_void_p), ("src_h", C.c_int32), ("src_w", C.c_int32), ("src_pitch", C.c_int32),
                ("rs_h", C.c_int32), ("rs_w", C.c_int32), ("top", C.c_int32), ("left", C.c_int32),
                ("pad_value", C.c_int32)]


def letterbox_geometry(h, w, new_shape=(640, 640), scale_fill=False):
    """Scalar part of LetterBox.__call__ (image_enhance/letter_box.py:35-58):
    -> dict(rs_w, rs_h, top, bottom, left, right, ratio=(rx, ry), dw, dh)."""
    if scale_fill:
        # cv2.resize(img, new_shape): dsize = (width, height) = new_shape; the reference then reads the ratio off
        # img.shape[0] / w and img.shape[1] / h (letter_box.py:44-45)
        rs_w, rs_h = int(new_shape[0]), int(new_shape[1])
        return dict(rs_w=rs_w, rs_h=rs_h, top=0, bottom=0, left=0, right=0, ratio=(rs_h / w, rs_w / h), dw=0, dh=0)
    ratio = (new_shape[0] / w, new_shape[1] / h)
    r = min(ratio)
    rs_w, rs_h = int(round(w * r)), int(round(h * r))
    dw, dh = (new_shape[0] - rs_w) / 2, (new_shape[1] - rs_h) / 2
    return dict(rs_w=rs_w, rs_h=rs_h, top=int(round(dh - 0.1)), bottom=int(round(dh + 0.1)),
                left=int(round(dw - 0.1)), right=int(round(dw + 0.1)), ratio=(r, r), dw=dw, dh=dh)


def letterbox_batch(images, new_shape=(640, 640), dtype=torch.float32, device=None, color=114, scale_fill=False,
                    out=None):
    """images: list of uint8 HWC (3-channel) arrays / tensors, host or device, any sizes.
    Returns (x [bs,3,H,W] `dtype` on the device, geometries) -- x is what
    np.concatenate([prepare_test_image(img)[0] for img in images]) holds in the reference."""
    if device is None:
        device = next((t.device for t in images if isinstance(t, torch.Tensor) and t.is_cuda), torch.device("cuda:0"))
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.YcError("letterbox_batch runs on a CUDA device: this path has no CPU fallback")
    if dtype not in (torch.float32, torch.bfloat16):
        raise _lib.YcError(f"unsupported output dtype {dtype}: use float32 or bfloat16")
    bs = len(images)
    geos, devs = [], []
    descs = (_LbImage * bs)()
    with torch.cuda.device(device):
        for i, img in enumerate(images):
            t = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img))
            if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
                raise _lib.YcError(f"image {i}: expected uint8 [H,W,3], got {t.dtype} {tuple(t.shape)}")
            t = t.to(device, non_blocking=True).contiguous()
            devs.append(t)
            h, w = int(t.shape[0]), int(t.shape[1])
            g = letterbox_geometry(h, w, new_shape, scale_fill)
            geos.append(g)
            d = descs[i]
            d.src, d.src_h, d.src_w, d.src_pitch = t.data_ptr(), h, w, 3 * w
            d.rs_h, d.rs_w, d.top, d.left, d.pad_value = g["rs_h"], g["rs_w"], g["top"], g["left"], int(color)
        g0 = geos[0]
        out_h, out_w = g0["top"] + g0["rs_h"] + g0["bottom"], g0["left"] + g0["rs_w"] + g0["right"]
        for i, g in enumerate(geos):
            if (g["top"] + g["rs_h"] + g["bottom"], g["left"] + g["rs_w"] + g["right"]) != (out_h, out_w):
                raise _lib.YcError(f"image {i}: letterboxed size differs from image 0 (non-square target with mixed "
                                   f"aspect ratios cannot be batched)")
        raw = np.frombuffer(descs, dtype=np.uint8)
        dd = torch.from_numpy(raw.copy()).to(device, non_blocking=True)
        if out is None:
            out = torch.empty((bs, 3, out_h, out_w), dtype=dtype, device=device)
        elif tuple(out.shape) != (bs, 3, out_h, out_w) or out.dtype != dtype or not out.is_contiguous():
            raise _lib.YcError("letterbox_batch: `out` has the wrong shape, dtype or layout")
        _lib.check(_lib.lib.yc_letterbox_batch(dd.data_ptr(), bs, out_h, out_w,
                                               _lib.YC_F32 if dtype == torch.float32 else _lib.YC_BF16, out.data_ptr(),
                                               _lib.stream_ptr(device)), "yc_letterbox_batch")
        for t in devs + [dd]:
            t.record_stream(torch.cuda.current_stream(device))
    return out, geos


class LetterBox(torch.nn.Module):
    """Reference signature (image_enhance/letter_box.py:10-27): LetterBox(new_shape, scale_fill_prob, color);
    __call__(img uint8 HWC, target_xyxy) -> (letterboxed uint8 HWC image, shifted targets)."""

    def __init__(self, new_shape=(640, 640), scale_fill_prob=1, color=(114, 114, 114)):
        super().__init__()
        self.new_shape = (new_shape, new_shape) if isinstance(new_shape, int) else new_shape
        self.color = color
        self.scale_fill_prob = scale_fill_prob

    def __call__(self, img, target_xyxy, device="cuda:0"):
        scale_fill = Random().random() < self.scale_fill_prob
        if len(set(self.color)) != 1:
            raise _lib.YcError("LetterBox: the device kernel pads with one grey level")
        x, geos = letterbox_batch([img], self.new_shape, torch.float32, device, self.color[0], scale_fill)
        g = geos[0]
        # the kernel's v / 255 is exact to invert: rint(x * 255) == v for every 8-bit v
        out = torch.round(x[0] * 255.0).to(torch.uint8).permute(1, 2, 0).contiguous().cpu().numpy()
        target_xyxy = np.copy(target_xyxy)
        target_xyxy[:, [0, 2]] = target_xyxy[..., [0, 2]] * g["ratio"][0] + g["dw"]
        target_xyxy[:, [1, 3]] = target_xyxy[..., [1, 3]] * g["ratio"][1] + g["dh"]
        return out, target_xyxy
