"""Box utilities (drop-in for reference utils/bbox.py) on the B200 kernels.

CUDA tensors go through the C ABI (`yc_cvt_bbox`, `yc_box_iou`, `yc_nms_*`); numpy arrays -- which
the reference also accepts in `cvt_bbox` -- are plain host data and are converted on the host.
"""
import math
from enum import Enum

import numpy as np
import torch

from .. import _lib
from .. import detect as _detect


class BBoxType(Enum):
    XYXY = 0
    XXYY = 2
    XYWH = 1


class CvtFlag(Enum):
    CVT_XXYY_XYXY = 0
    CVT_XXYY_XYWH = 1
    CVT_XYXY_XXYY = 2
    CVT_XYXY_XYWH = 3
    CVT_XYWH_XXYY = 4
    CVT_XYWH_XYXY = 5


def check(flag: CvtFlag):
    return flag.value in range(6)


def _cvt_host(b, f):
    """Column algebra of reference utils/bbox.py:36-57 on a host array (result starts as a copy)."""
    r = b.copy()
    c0, c1, c2, c3 = (b[:, i] for i in range(4))
    if f in (0, 2):
        r[:, 1], r[:, 2] = c2, c1
    elif f == 1:
        r[:, 2], r[:, 3] = c1 - c0, c3 - c2
        r[:, 0], r[:, 1] = c0 + r[:, 2] / 2, c2 + r[:, 3] / 2
    elif f == 3:
        r[:, 2:4] = b[:, 2:4] - b[:, 0:2]
        r[:, 0:2] = b[:, 0:2] + r[:, 2:4] / 2
    elif f == 4:
        r[:, 0], r[:, 1] = c0 - c2 / 2, c0 + c2 / 2
        r[:, 2], r[:, 3] = c1 - c3 / 2, c1 + c3 / 2
    elif f == 5:
        r[:, 0], r[:, 1] = c0 - c2 / 2, c1 - c3 / 2
        r[:, 2], r[:, 3] = c0 + c2 / 2, c1 + c3 / 2
    return r


def cvt_bbox(bbox, flag: CvtFlag):
    """Layout conversion between XXYY / XYXY / XYWH (reference utils/bbox.py:29-59)."""
    if not check(flag):
        raise Exception()
    if isinstance(bbox, np.ndarray):
        return _cvt_host(bbox, flag.value)
    if not bbox.is_cuda:
        raise _lib.YcError("cvt_bbox: tensors must live on a CUDA device (no CPU fallback); pass a numpy array "
                           "for host data")
    if bbox.dtype != torch.float32 or bbox.dim() != 2 or bbox.shape[1] != 4:
        raise _lib.YcError("cvt_bbox: expected a float32 [n,4] tensor")
    src = bbox.contiguous()
    out = torch.empty_like(src)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib.yc_cvt_bbox(src.data_ptr(), src.shape[0], flag.value, out.data_ptr(),
                                        _lib.stream_ptr(src.device)), "yc_cvt_bbox")
    return out


def box_iou(box1, box2):
    """Pairwise IoU [N,M] of xyxy boxes (reference utils/bbox.py:62-72)."""
    _lib.require_cuda(box1, "box1")
    _lib.require_cuda(box2, "box2")
    b1, b2 = box1.float().contiguous(), box2.float().contiguous()
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    with torch.cuda.device(b1.device):
        _lib.check(_lib.lib.yc_box_iou(b1.data_ptr(), b1.shape[0], b2.data_ptr(), b2.shape[0], out.data_ptr(),
                                       _lib.stream_ptr(b1.device)), "yc_box_iou")
    return out


def bbox_iou(box1, box2, x1y1x2y2=True, giou=False, diou=False, ciou=False, eps=1e-7):
    """IoU / GIoU / DIoU / CIoU of one box against n boxes (reference utils/bbox.py:75-118).

    Training-loss utility (losses/yolo_loss.py); not on the inference hot path, so it is plain
    tensor algebra on whatever device the inputs live on.
    """
    box2 = box2.T
    if x1y1x2y2:
        ax1, ay1, ax2, ay2 = box1[0], box1[1], box1[2], box1[3]
        bx1, by1, bx2, by2 = box2[0], box2[1], box2[2], box2[3]
    else:
        ax1, ax2 = box1[0] - box1[2] / 2, box1[0] + box1[2] / 2
        ay1, ay2 = box1[1] - box1[3] / 2, box1[1] + box1[3] / 2
        bx1, bx2 = box2[0] - box2[2] / 2, box2[0] + box2[2] / 2
        by1, by2 = box2[1] - box2[3] / 2, box2[1] + box2[3] / 2
    iw = (torch.min(ax2, bx2) - torch.max(ax1, bx1)).clamp(0)
    ih = (torch.min(ay2, by2) - torch.max(ay1, by1)).clamp(0)
    inter = iw * ih
    w1, h1 = ax2 - ax1, ay2 - ay1 + eps
    w2, h2 = bx2 - bx1, by2 - by1 + eps
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    if not (giou or diou or ciou):
        return iou
    cw = torch.max(ax2, bx2) - torch.min(ax1, bx1)
    ch = torch.max(ay2, by2) - torch.min(ay1, by1)
    if giou and not (diou or ciou):
        c_area = cw * ch + eps
        return iou - (c_area - union) / c_area
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = ((bx1 + bx2 - ax1 - ax2) ** 2 + (by1 + by2 - ay1 - ay2) ** 2) / 4
    if diou:
        return iou - rho2 / c2
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


def non_max_suppression(self, prediction, num_classes, input_shape, image_shape, letterbox_image,
                        conf_thres=0.5, nms_thres=0.4):
    """Twin of detect.non_max_suppression with the reference's stray leading `self` parameter
    (utils/bbox.py:121-128); `self` is ignored."""
    return _detect.non_max_suppression(prediction, num_classes, input_shape, image_shape, letterbox_image,
                                       conf_thres=conf_thres, nms_thres=nms_thres)


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms drop-in (the module-global `nms` of reference utils/bbox.py:7 and detect.py:6)."""
    return _detect.nms(boxes, scores, iou_threshold)


def make_grid(nx=20, ny=20):
    """Cell-index grid (1,1,ny,nx,2) with [...,0]=x, [...,1]=y (reference utils/bbox.py:201-204)."""
    yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
    return torch.stack((xv, yv), 2).view((1, 1, ny, nx, 2)).float()
