"""Decode + NMS entry points (drop-in for the post-backbone functions of reference detect.py).

`decode_box`, `non_max_suppression` and `yolo_correct_boxes` keep the reference signatures
(detect.py:29, :90, :147).  `non_max_suppression` runs thresholding, compaction, per-class NMS and
the letterbox undo for the whole batch on the device (one C-ABI call, `yc_nms_batched`) and does a
single device->host read at the end to honour the ndarray return type.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_WS = {}


def _workspace(device, nbytes):
    key = (device.type, device.index)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def decode_box(inputs, anchors, anchors_mask, num_labels, image_size=(640, 640)):
    """Variant A decode (reference detect.py:29-87): list of raw conv maps [bs, na*(5+nc), H, W]
    -> list of [bs, na*H*W, 5+nc] with normalised xywh and sigmoid scores."""
    anchors = np.asarray(anchors, dtype=np.float64).reshape(-1, 2)
    no = num_labels + 5
    outs = []
    for i, pred in enumerate(inputs):
        _lib.require_cuda(pred, "decode_box input")
        if pred.dtype != torch.float32:
            raise _lib.YcError("decode_box: float32 conv maps expected")
        pred = pred.contiguous()
        bs, ch, ny, nx = pred.shape
        na = len(anchors_mask[i])
        if ch != na * no:
            raise _lib.YcError(f"decode_box: level {i} has {ch} channels, expected {na * no}")
        stride_h, stride_w = image_size[0] / ny, image_size[0] / nx  # reference detect.py:38-39
        scaled = (C.c_float * (na * 2))()
        for j, (aw, ah) in enumerate(anchors[anchors_mask[i]]):
            scaled[2 * j], scaled[2 * j + 1] = aw / stride_w, ah / stride_h
        out = torch.empty((bs, na * ny * nx, no), dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(_lib.lib.yc_decode_box(pred.data_ptr(), bs, na, no, ny, nx, scaled, out.data_ptr(),
                                              _lib.stream_ptr(pred.device)), "yc_decode_box")
        outs.append(out)
    return outs


def nms_device(prediction, num_classes, conf_thres, nms_thres, input_shape=None, image_shape=None,
               letterbox_image=False, write_corners=True, box_div=None):
    """Batched threshold + per-class NMS on the device, results left on the device.

    Returns (rows [total,7], idx [total] original row per detection, counts [bs], offsets [bs+1]),
    all CUDA tensors; rows/idx are views of a capacity-sized buffer (valid up to offsets[bs]).
    With image_shape given the rows hold y1,x1,y2,x2 in image pixels (yolo_correct_boxes), else
    x1,y1,x2,y2 as decoded.  box_div = (w, h): cx,w /= w and cy,h /= h (IEEE division) before the corners are formed
    (`box_div_w/h` of yc_nms_params: IDetect's input-pixel boxes -> the normalised boxes yolo_correct_boxes expects).
    """
    _lib.require_cuda(prediction, "prediction")
    if prediction.dtype != torch.float32 or prediction.dim() != 3 or not prediction.is_contiguous():
        raise _lib.YcError("prediction must be a contiguous float32 [bs, rows, 5+nc] tensor")
    bs, rows, stride = prediction.shape
    dev = prediction.device
    p = _lib.NmsParams()
    p.bs, p.rows, p.row_stride, p.nc = bs, rows, stride, num_classes
    p.conf_thres, p.nms_thres = float(conf_thres), float(nms_thres)
    p.write_corners = 1 if write_corners else 0
    if box_div is not None:
        p.box_div_w, p.box_div_h = float(box_div[0]), float(box_div[1])
    hw = None
    if image_shape is not None:
        p.correct_boxes, p.letterbox = 1, 1 if letterbox_image else 0
        p.input_h, p.input_w = int(input_shape[0]), int(input_shape[1])
        hw_np = np.asarray(image_shape, dtype=np.int32).reshape(-1, 2)
        hw = torch.from_numpy(np.ascontiguousarray(hw_np)).to(dev)
        p.image_hw, p.image_hw_stride = hw.data_ptr(), (2 if hw_np.shape[0] > 1 else 0)
        if hw_np.shape[0] not in (1, bs):
            raise _lib.YcError("image_shape must be one (h, w) or one per image")
    with torch.cuda.device(dev):
        ws = _workspace(dev, _lib.lib.yc_nms_workspace_bytes(bs, rows, num_classes))
        out_rows = torch.empty((bs * rows, 7), dtype=torch.float32, device=dev)
        out_idx = torch.empty((bs * rows,), dtype=torch.int32, device=dev)
        counts = torch.empty((bs,), dtype=torch.int32, device=dev)
        offsets = torch.empty((bs + 1,), dtype=torch.int32, device=dev)
        _lib.check(_lib.lib.yc_nms_batched(prediction.data_ptr(), C.byref(p), ws.data_ptr(), ws.numel(),
                                           out_rows.data_ptr(), out_idx.data_ptr(), counts.data_ptr(),
                                           offsets.data_ptr(), _lib.stream_ptr(dev)), "yc_nms_batched")
    return out_rows, out_idx, counts, offsets


def non_max_suppression(prediction, num_classes, input_shape, image_shape, letterbox_image,
                        conf_thres=0.5, nms_thres=0.4, return_indices=False, box_div=None):
    """Reference detect.py:90-144.  prediction [bs, rows, 5+nc] (xywh + obj + cls) on a CUDA device;
    its first four columns are overwritten with corners, as the reference does (detect.py:103).
    Returns a list with, per image, None or ndarray[n,7] = y1,x1,y2,x2 (image px), obj, class_conf,
    class_id -- classes ascending, score descending within a class."""
    rows, idx, counts, offsets = nms_device(prediction, num_classes, conf_thres, nms_thres, input_shape,
                                            image_shape, letterbox_image, write_corners=True, box_div=box_div)
    off = offsets.cpu().numpy()
    total = int(off[-1])
    host = rows[:total].cpu().numpy()
    out = [None if off[b + 1] == off[b] else host[off[b]:off[b + 1]].copy() for b in range(len(off) - 1)]
    if return_indices:
        hidx = idx[:total].cpu().numpy()
        return out, [hidx[off[b]:off[b + 1]].astype(np.int64) for b in range(len(off) - 1)]
    return out


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms drop-in (call site reference detect.py:133): int64 indices, score order."""
    _lib.require_cuda(boxes, "boxes")
    n = boxes.shape[0]
    dev = boxes.device
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=dev)
    b, s = boxes.float().contiguous(), scores.float().contiguous()
    with torch.cuda.device(dev):
        ws = _workspace(dev, _lib.lib.yc_nms_workspace_bytes(1, n, 1))
        keep = torch.empty((n,), dtype=torch.int32, device=dev)
        cnt = torch.empty((1,), dtype=torch.int32, device=dev)
        _lib.check(_lib.lib.yc_nms_single(b.data_ptr(), s.data_ptr(), n, float(iou_threshold), ws.data_ptr(),
                                          ws.numel(), keep.data_ptr(), cnt.data_ptr(), _lib.stream_ptr(dev)),
                   "yc_nms_single")
    return keep[:int(cnt.item())].long()


def yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image):
    """Letterbox undo on host arrays (reference detect.py:147-165): normalised centre/size ->
    [y1, x1, y2, x2] in original-image pixels.  The batched device version lives in the gather
    step of `yc_nms_batched`; this host form is kept for callers that hold numpy boxes."""
    yx, hw = box_xy[..., ::-1], box_wh[..., ::-1]
    input_shape, image_shape = np.array(input_shape), np.array(image_shape)
    if letterbox_image:
        fitted = np.round(image_shape * np.min(input_shape / image_shape))
        yx = (yx - (input_shape - fitted) / 2. / input_shape) * (input_shape / fitted)
        hw *= input_shape / fitted
    lo, hi = yx - hw / 2., yx + hw / 2.
    boxes = np.concatenate([lo[..., 0:1], lo[..., 1:2], hi[..., 0:1], hi[..., 1:2]], axis=-1)
    boxes *= np.concatenate([image_shape, image_shape], axis=-1)
    return boxes


def prepare_test_images(images, target_size, dtype=torch.float32, device="cuda:0"):
    """Batched prepare_test_image (reference detect.py:16-26) for already decoded images (uint8 HWC, BGR as
    cv2.imread returns them): letterbox + /255 + CHW on the device.  Returns (x [bs,3,H,W], original images)."""
    from .image_enhance import letterbox_batch
    x, _ = letterbox_batch(images, target_size, dtype, device)
    return x, images


def format_detections(rows, offsets, image_hw):
    """The formatting loop of reference predict (detect.py:236-258) for a whole batch on the device.
    rows [>=total,7] (y1,x1,y2,x2,obj,class_conf,class) and offsets [bs+1] as nms_device / PostBackbone return
    them; image_hw int32 [bs,2] or [2] (shared).  Returns device tensors (box_xyxy int32 [cap,4], conf float32 [cap],
    label int32 [cap]) valid up to offsets[bs]."""
    _lib.require_cuda(rows, "rows")
    dev = rows.device
    bs = offsets.numel() - 1
    hw = torch.as_tensor(image_hw, dtype=torch.int32).reshape(-1, 2).to(dev).contiguous()
    if hw.shape[0] not in (1, bs):
        raise _lib.YcError("format_detections: image_hw must hold one or bs (h, w) pairs")
    cap = rows.shape[0]
    box = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    conf = torch.empty((cap,), dtype=torch.float32, device=dev)
    label = torch.empty((cap,), dtype=torch.int32, device=dev)
    off = offsets.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.yc_format_detections(rows.contiguous().data_ptr(), off.data_ptr(), bs, hw.data_ptr(),
                                                 2 if hw.shape[0] > 1 else 0, box.data_ptr(), conf.data_ptr(),
                                                 label.data_ptr(), _lib.stream_ptr(dev)), "yc_format_detections")
    return box, conf, label


def detect_post_backbone(head, features, input_shape, image_shape, letterbox_image=True, conf_thres=0.3,
                         nms_thres=0.3):
    """The post-backbone half of reference detect.predict (detect.py:227-234) for an I*Detect head:
    features (list of neck maps) -> head forward (decode fused) -> batched NMS -> per-image arrays.
    IDetect boxes are in input pixels; they are normalised by the input size before NMS so that
    yolo_correct_boxes sees what decode_box would give it."""
    was = head.return_raw
    head.return_raw = False
    try:
        z, _ = head(list(features))
    finally:
        head.return_raw = was
    return non_max_suppression(z, head.nc, input_shape, image_shape, letterbox_image, conf_thres, nms_thres,
                               box_div=(input_shape[1], input_shape[0]))
