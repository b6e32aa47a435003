"""torch restatement of the reference's post-backbone path (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference is pure Python over torch + torchvision, so its own CPU path *is* a sequence of torch
CPU ops.  /root/reference does not exist on the GPU box, therefore bench.py's `cpu_baseline` and
`--impl reference` legs time this port (kind "port"): the same torch ops in the same order as
  IDetect.forward            nets/idetect.py:26-45   (ImplicitA/M: nets/common.py:425-426,438-439)
  detect.non_max_suppression detect.py:90-144        (torchvision.ops.nms at :133)
  detect.yolo_correct_boxes  detect.py:147-165
written from the survey's description, not copied.  tests/test_oracle_golden.py pins it against the
fixtures generated from the unmodified reference (bit-exact: same ops, same library).
The ops are device-agnostic, as the reference's are: fed CUDA tensors it runs cuDNN/cuBLAS convolutions, ~50 elementwise
launches per forward and torchvision's CUDA nms kernels with the reference's per-image / per-class Python loops and
host round trips (detect.py:124 `.cpu().unique()`, :140 `.cpu().numpy()`) -- bench.py's `library_baseline`.
Never imported by the product package.
"""
import numpy as np
import torch
from torchvision.ops import nms as tv_nms


def idetect_forward(p, xs, strides):
    """p: dict with per-level 'w' [N,K], 'b', 'ia', 'im' (numpy or tensors) and 'anchors' [nl,na,2]."""
    dev = torch.as_tensor(xs[0]).device
    anchors = torch.as_tensor(np.asarray(p["anchors"]), dtype=torch.float32).to(dev)
    nl, na = anchors.shape[0], anchors.shape[1]
    zs, raws = [], []
    for i in range(nl):
        x = torch.as_tensor(xs[i])
        w = torch.as_tensor(p["w"][i]).reshape(-1, x.shape[1], 1, 1)
        t = torch.nn.functional.conv2d(torch.as_tensor(p["ia"][i]).reshape(1, -1, 1, 1) + x, w, torch.as_tensor(p["b"][i]))
        t = torch.as_tensor(p["im"][i]).reshape(1, -1, 1, 1) * t
        bs, _, ny, nx = t.shape
        no = t.shape[1] // na
        t = t.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        raws.append(t)
        gy, gx = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
        grid = torch.stack((gx, gy), 2).view(1, 1, ny, nx, 2).float().to(dev)   # nets/idetect.py:38 `.to(x[i].device)`
        y = t.sigmoid()
        y[..., 0:2] = (y[..., 0:2] * 2. - 0.5 + grid) * float(strides[i])
        y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * anchors[i].view(1, na, 1, 1, 2)
        zs.append(y.view(bs, -1, no))
    return torch.cat(zs, 1), raws


def correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image):
    yx, hw = box_xy[..., ::-1], box_wh[..., ::-1]
    input_shape, image_shape = np.array(input_shape), np.array(image_shape)
    if letterbox_image:
        new = np.round(image_shape * np.min(input_shape / image_shape))
        yx = (yx - (input_shape - new) / 2. / input_shape) * (input_shape / new)
        hw *= input_shape / new
    lo, hi = yx - hw / 2., yx + hw / 2.
    out = np.concatenate([lo[..., 0:1], lo[..., 1:2], hi[..., 0:1], hi[..., 1:2]], axis=-1)
    out *= np.concatenate([image_shape, image_shape], axis=-1)
    return out


def non_max_suppression(prediction, num_classes, input_shape, image_shape, letterbox_image, conf_thres=0.5,
                        nms_thres=0.4):
    """prediction: float32 tensor [bs, rows, 5+nc] (xywh); modified in place like the reference."""
    corner = prediction.new(prediction.shape)
    corner[:, :, 0] = prediction[:, :, 0] - prediction[:, :, 2] / 2
    corner[:, :, 1] = prediction[:, :, 1] - prediction[:, :, 3] / 2
    corner[:, :, 2] = prediction[:, :, 0] + prediction[:, :, 2] / 2
    corner[:, :, 3] = prediction[:, :, 1] + prediction[:, :, 3] / 2
    prediction[:, :, :4] = corner[:, :, :4]
    out = [None] * len(prediction)
    for i, img in enumerate(prediction):
        conf, cls = torch.max(img[:, 5:5 + num_classes], 1, keepdim=True)
        keep = (img[:, 4] * conf[:, 0] >= conf_thres).squeeze()
        img, conf, cls = img[keep], conf[keep], cls[keep]
        if not img.size(0):
            continue
        det = torch.cat((img[:, :5], conf.float(), cls.float()), 1)
        for c in det[:, -1].cpu().unique():                  # detect.py:124
            dc = det[det[:, -1] == c]
            k = tv_nms(dc[:, :4], dc[:, 4] * dc[:, 5], nms_thres)
            out[i] = dc[k] if out[i] is None else torch.cat((out[i], dc[k]))
        if out[i] is not None:
            o = out[i].cpu().numpy()                         # detect.py:140
            xy, wh = (o[:, 0:2] + o[:, 2:4]) / 2, o[:, 2:4] - o[:, 0:2]
            o[:, :4] = correct_boxes(xy, wh, input_shape, image_shape, letterbox_image)
            out[i] = o
    return out


def post_backbone(p, xs, strides, nc, input_shape, image_shape, letterbox_image, conf_thres, nms_thres):
    """features -> detections, the order of detect.predict (detect.py:227-234) with an IDetect head."""
    z, _ = idetect_forward(p, xs, strides)
    z[..., 0] /= input_shape[1]; z[..., 2] /= input_shape[1]
    z[..., 1] /= input_shape[0]; z[..., 3] /= input_shape[0]
    return non_max_suppression(z, nc, input_shape, image_shape, letterbox_image, conf_thres, nms_thres)
