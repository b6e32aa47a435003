"""ctypes/numpy front-end of the C oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.  Function names follow
the reference's (nets/idetect.py, nets/iaux_detect.py, nets/ibin.py, detect.py,
utils/bbox.py); each wrapper states which C function (and thus which reference
file:line) it drives.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = _build.LIB if os.path.exists(_build.LIB) and os.path.getmtime(
            _build.LIB) >= os.path.getmtime(_build.SRC) else _build.build()
        _LIB = C.CDLL(path)
        _LIB.yco_nms.restype = C.c_int
        _LIB.yco_nms_image.restype = C.c_int
        _LIB.yco_cvt_bbox.restype = C.c_int
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def head_level(x, w, b, ia, im, na, no):
    """yco_head_level: ImplicitA -> 1x1 conv -> ImplicitM -> [bs,na,ny,nx,no] raw logits."""
    x = _f32(x)
    bs, K, H, W = x.shape
    w = _f32(w).reshape(na * no, K)
    b = _f32(b) if b is not None else None
    ia = _f32(ia).reshape(-1) if ia is not None else None
    im = _f32(im).reshape(-1) if im is not None else None
    raw = np.empty((bs, na, H, W, no), np.float32)
    lib().yco_head_level(_p(x), _p(w), _p(b), _p(ia), _p(im), C.c_int(bs), C.c_int(K),
                         C.c_int(H * W), C.c_int(na), C.c_int(no), _p(raw))
    return raw


def head_forward(kind, params, xs, strides):
    """Inference forward of IDetect / IAuxDetect / IBin (nets/*.py forward, eval mode).

    params: dict with lists 'w','b','ia','im' (per level), 'anchors' [nl,na,2] (pixels),
            for 'iaux' also 'w2','b2'; for 'ibin' also 'bins_w','bins_h' and 'bin_count'.
    xs:     list of [bs,K,H,W] arrays (nl of them; 2*nl for 'iaux').
    Returns (z [bs, sum(na*H*W), nc+5], raws list of [bs,na,ny,nx,no]).
    For 'iaux' the aux raws are returned as a third value (the reference computes and
    then discards them, nets/iaux_detect.py:37-38,49).
    """
    anchors = _f32(params["anchors"])
    nl, na = anchors.shape[0], anchors.shape[1]
    bs = xs[0].shape[0]
    no = params["w"][0].shape[0] // na
    raws = [head_level(xs[i], params["w"][i], params["b"][i], params["ia"][i],
                       params["im"][i], na, no) for i in range(nl)]
    rows = sum(na * r.shape[2] * r.shape[3] for r in raws)
    if kind == "ibin":
        len_ = params["bin_count"] + 1
        no_out = no - 2 * len_ + 2
    else:
        no_out = no
    z = np.empty((bs, rows, no_out), np.float32)
    off = 0
    for i, r in enumerate(raws):
        ny, nx = r.shape[2], r.shape[3]
        a_wh = _f32(anchors[i].reshape(-1))
        if kind == "ibin":
            bw, bh = _f32(params["bins_w"]), _f32(params["bins_h"])
            assert np.array_equal(bw, bh)
            step = np.float32(4.0 / params["bin_count"])
            lib().yco_decode_ibin(_p(r), C.c_int(bs), C.c_int(na), C.c_int(ny), C.c_int(nx),
                                  C.c_int(no), C.c_int(params["bin_count"]),
                                  C.c_float(float(strides[i])), _p(a_wh), _p(bw),
                                  C.c_float(2.0), C.c_float(float(step)), C.c_float(0.0),
                                  C.c_float(4.0), _p(z), C.c_int(rows), C.c_int(off))
        else:
            lib().yco_decode_idetect(_p(r), C.c_int(bs), C.c_int(na), C.c_int(ny),
                                     C.c_int(nx), C.c_int(no), C.c_float(float(strides[i])),
                                     _p(a_wh), _p(z), C.c_int(rows), C.c_int(off))
        off += na * ny * nx
    if kind == "iaux":
        aux = [head_level(xs[i + nl], params["w2"][i], params["b2"][i], None, None, na, no)
               for i in range(nl)]
        return z, raws, aux
    return z, raws


def decode_box(inputs, anchors, anchors_mask, num_labels, image_size=(640, 640)):
    """detect.decode_box (detect.py:29-87) through yco_decode_box_level."""
    anchors = np.asarray(anchors, dtype=np.float64).reshape(-1, 2)
    outs = []
    for i, pred in enumerate(inputs):
        pred = _f32(pred)
        bs, _, ny, nx = pred.shape
        sel = np.ascontiguousarray(anchors[anchors_mask[i]].reshape(-1))
        na = len(anchors_mask[i])
        no = num_labels + 5
        out = np.empty((bs, na * ny * nx, no), np.float32)
        lib().yco_decode_box_level(_p(pred), C.c_int(bs), C.c_int(na), C.c_int(ny),
                                   C.c_int(nx), C.c_int(no), _p(sel, C.c_double),
                                   C.c_double(float(image_size[0])), _p(out))
        outs.append(out)
    return outs


def nms(boxes, scores, thr):
    """torchvision.ops.nms restatement (yco_nms). Returns int64 keep indices."""
    boxes, scores = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    keep = np.empty(max(len(scores), 1), np.int32)
    n = lib().yco_nms(_p(boxes), _p(scores), C.c_int(len(scores)), C.c_double(float(thr)),
                      _p(keep, C.c_int))
    return keep[:n].astype(np.int64)


def nms_candidates(prediction, num_classes, conf_thres=0.5, nms_thres=0.4):
    """Device part of detect.non_max_suppression (detect.py:97-137), per image.

    prediction [bs, n, 5+nc] float32 C-contiguous is modified in place (corners), as in
    the reference.  Returns (rows list of ndarray[k,7] xyxy, idx list of ndarray[k]).
    """
    assert prediction.dtype == np.float32 and prediction.flags.c_contiguous
    bs, n, no = prediction.shape
    assert no >= 5 + num_classes
    rows, idxs = [], []
    for b in range(bs):
        if no != 5 + num_classes:
            sub = np.ascontiguousarray(prediction[b][:, :5 + num_classes])
        else:
            sub = prediction[b]
        o = np.empty((n, 7), np.float32)
        ix = np.empty(max(n, 1), np.int32)
        k = lib().yco_nms_image(_p(sub), C.c_int(n), C.c_int(num_classes),
                                C.c_float(float(np.float32(conf_thres))),
                                C.c_double(float(nms_thres)), _p(o), _p(ix, C.c_int))
        if sub is not prediction[b]:
            prediction[b][:, :4] = sub[:, :4]
        rows.append(o[:k].copy())
        idxs.append(ix[:k].copy())
    return rows, idxs


def correct_boxes_rows(rows, input_shape, image_shape, letterbox_image):
    """detect.py:140-142 + yolo_correct_boxes (detect.py:147-165) on [k,7] rows, in place."""
    rows = np.ascontiguousarray(rows, np.float32)
    lib().yco_correct_boxes(_p(rows), C.c_int(rows.shape[0]), C.c_int(int(input_shape[0])),
                            C.c_int(int(input_shape[1])), C.c_int(int(image_shape[0])),
                            C.c_int(int(image_shape[1])), C.c_int(1 if letterbox_image else 0))
    return rows


def non_max_suppression(prediction, num_classes, input_shape, image_shape, letterbox_image,
                        conf_thres=0.5, nms_thres=0.4, return_indices=False):
    """detect.non_max_suppression (detect.py:90-144): list of None | ndarray[k,7] (yxyx px)."""
    rows, idxs = nms_candidates(prediction, num_classes, conf_thres, nms_thres)
    out = [None if r.shape[0] == 0 else
           correct_boxes_rows(r, input_shape, image_shape, letterbox_image) for r in rows]
    return (out, idxs) if return_indices else out


def box_iou(b1, b2):
    b1, b2 = _f32(b1).reshape(-1, 4), _f32(b2).reshape(-1, 4)
    out = np.empty((b1.shape[0], b2.shape[0]), np.float32)
    lib().yco_box_iou(_p(b1), C.c_int(b1.shape[0]), _p(b2), C.c_int(b2.shape[0]), _p(out))
    return out


def cvt_bbox(bbox, flag):
    bbox = _f32(bbox).reshape(-1, 4)
    out = np.empty_like(bbox)
    rc = lib().yco_cvt_bbox(_p(bbox), C.c_int(bbox.shape[0]), C.c_int(int(flag)), _p(out))
    if rc != 0:
        raise Exception()
    return out


# ---- preprocessing / formatting either side of the model ---------------------------------
def letterbox_geometry(h, w, new_shape=(640, 640)):
    """Scalar part of LetterBox.__call__ without scale_fill (image_enhance/letter_box.py:38-58):
    -> (rs_w, rs_h, top, bottom, left, right, ratio, dw, dh)."""
    ratio = (new_shape[0] / w, new_shape[1] / h)
    r = min(ratio)
    rs_w, rs_h = int(round(w * r)), int(round(h * r))
    dw, dh = (new_shape[0] - rs_w) / 2, (new_shape[1] - rs_h) / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return rs_w, rs_h, top, bottom, left, right, r, dw, dh


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HWC images: OpenCV's fixed-point
    bilinear (third-party, opencv-python 4.13, imgproc/resize.cpp: 11-bit coefficients, int32 horizontal pass,
    vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2; call site image_enhance/letter_box.py:53)."""
    sh, sw = src.shape[:2]

    def coefs(ssize, dsize, clamp):
        scale = 1.0 / (dsize / ssize)
        idx, a = np.zeros(dsize, np.int64), np.zeros((dsize, 2), np.int64)
        for d in range(dsize):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(np.floor(f))
            f = np.float32(f - np.float32(s))
            if clamp:
                if s < 0:
                    f, s = np.float32(0), 0
                if s >= ssize - 1:
                    f, s = np.float32(0), ssize - 1
            idx[d] = s
            a[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
            a[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
        return idx, a

    xi, xa = coefs(sw, dw, True)
    yi, ya = coefs(sh, dh, False)
    s = src.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    rows = s[:, xi] * xa[:, 0][None, :, None] + s[:, x1] * xa[:, 1][None, :, None]
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def prepare_test_image(image, target_size, color=114):
    """prepare_test_image (detect.py:16-26) on a decoded uint8 HWC image: LetterBox(target_size, scale_fill_prob=0)
    -> float32 / 255 -> CHW -> [1,3,H,W]."""
    h, w = image.shape[:2]
    rs_w, rs_h, top, bottom, left, right, _, _, _ = letterbox_geometry(h, w, target_size)
    img = resize_linear_u8(image, rs_w, rs_h)   # the reference skips the call when the size is unchanged: identity
    out = np.full((top + rs_h + bottom, left + rs_w + right, 3), color, np.uint8)
    out[top:top + rs_h, left:left + rs_w] = img
    return np.expand_dims(np.transpose(np.array(out, dtype='float32') / 255., (2, 0, 1)), 0)


def format_detections(rows, image_shape):
    """The formatting loop of predict (detect.py:236-258) for one image: rows [n,7] (y1,x1,y2,x2,obj,cls_conf,cls)
    -> (box_xyxy int32 [n,4], conf float32 [n], label int32 [n])."""
    rows = np.asarray(rows, np.float32).reshape(-1, 7)
    label = np.array(rows[:, 6], dtype='int32')
    conf = rows[:, 4] * rows[:, 5]
    box = np.empty((rows.shape[0], 4), np.int32)
    for i in range(rows.shape[0]):
        y1, x1, y2, x2 = rows[i, :4]
        box[i] = [max(0, np.floor(x1).astype('int32')), max(0, np.floor(y1).astype('int32')),
                  min(image_shape[1], np.floor(x2).astype('int32')), min(image_shape[0], np.floor(y2).astype('int32'))]
    return box, conf, label


def evaluate_map(dets, gts, num_classes, iou_thresholds):
    """CPU restatement of the evaluator (yolo_continuous_b200/evaluate.py; the reference has no mAP code -- "parity
    unpinned" for this component: the definition is the usual one, stated here).
    dets: list over images of None | ndarray[n,7] (box[4], obj, class_conf, class) in the order NMS leaves them (class
    ascending, score descending); gts: list over images of (boxes [m,4], labels [m]).  For every IoU threshold each
    detection, in that order, takes the unmatched ground-truth box of its class with the highest IoU >= thr (ties: the
    first); IoU = utils/bbox.py:62-72 in binary32.  AP per class = mean over the 101 recall points 0, .01, ... 1 of the
    precision envelope (0 beyond the reached recall).  Returns (tp list over images of [T, n] uint8, ap [T, nc] float64
    with NaN for classes without ground truth)."""
    T = len(iou_thresholds)
    tps, scores, classes = [], [], []
    n_gt = np.zeros(num_classes, np.int64)
    for det, (gb, gl) in zip(dets, gts):
        gb, gl = np.asarray(gb, np.float32).reshape(-1, 4), np.asarray(gl).reshape(-1)
        for c in gl:
            n_gt[int(c)] += 1
        if det is None or len(det) == 0:
            tps.append(np.zeros((T, 0), np.uint8))
            continue
        iou = box_iou(np.ascontiguousarray(det[:, :4], np.float32), gb) if len(gb) else np.zeros((len(det), 0), np.float32)
        tp = np.zeros((T, len(det)), np.uint8)
        for t, thr in enumerate(iou_thresholds):
            taken = np.zeros(len(gb), bool)
            for d in range(len(det)):
                best, bi = -1.0, -1
                for g in range(len(gb)):
                    if int(gl[g]) != int(det[d, 6]) or taken[g]:
                        continue
                    v = iou[d, g]
                    if v >= np.float32(thr) and v > best:
                        best, bi = v, g
                if bi >= 0:
                    taken[bi] = True
                    tp[t, d] = 1
        tps.append(tp)
        scores.append((det[:, 4] * det[:, 5]).astype(np.float32))
        classes.append(det[:, 6].astype(np.int64))
    ap = np.full((T, num_classes), np.nan)
    ap[:, n_gt > 0] = 0.0
    if scores:
        score, cls, tp = np.concatenate(scores), np.concatenate(classes), np.concatenate([t for t in tps if t.shape[1]], 1)
        rec_thr = np.linspace(0, 1, 101)
        for c in range(num_classes):
            m = np.nonzero(cls == c)[0]
            if n_gt[c] == 0 or len(m) == 0:
                continue
            o = m[np.argsort(-score[m], kind="stable")]
            for t in range(T):
                ctp = np.cumsum(tp[t, o].astype(np.float64))
                recall, prec = ctp / n_gt[c], ctp / np.arange(1, len(o) + 1)
                env = np.maximum.accumulate(prec[::-1])[::-1]
                idx = np.searchsorted(recall, rec_thr, side="left")
                ap[t, c] = np.where(idx < len(o), env[np.minimum(idx, len(o) - 1)], 0.0).mean()
    return tps, ap
