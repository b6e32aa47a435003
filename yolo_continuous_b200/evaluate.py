"""Batched on-device evaluator for the detections of the post-backbone path.

The reference has no evaluator (no mAP code anywhere); this one consumes exactly what
`PostBackbone` / `nms_device` leave on the device -- rows [total,7] = (y1,x1,y2,x2,obj,class_conf,class) with per-image
offsets -- so the mAP-eval configuration (conf 0.001 / iou 0.65, millions of detections per step) never leaves the GPU.
Matching runs in `yc_match_detections` (one warp per image and IoU threshold, IoU = the reference's box_iou,
utils/bbox.py:62-72); the per-class precision/recall integration (101-point interpolated AP, the COCO convention) is a
handful of sorts and prefix sums on device tensors.  Definition restated on the CPU in oracle/oracle.py (evaluate_map).
"""
import ctypes as C

import torch

from . import _lib


class DetectionEvaluator:
    def __init__(self, num_classes, iou_thresholds=(0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95), device="cuda:0"):
        self.nc = int(num_classes)
        self.device = torch.device(device)
        self.thrs = torch.tensor(list(iou_thresholds), dtype=torch.float32, device=self.device)
        self.reset()

    def reset(self):
        self._score, self._cls, self._tp = [], [], []
        self._n_gt = torch.zeros(self.nc, dtype=torch.int64, device=self.device)

    def update(self, rows, offsets, gt_boxes, gt_labels, gt_offsets):
        """rows [>=total,7] and offsets [bs+1] (int32) of a batch as the pipeline returns them; gt_boxes [n_gt,4] float32
        in the same coordinate convention as rows[:, :4] (y1,x1,y2,x2 image pixels after the letterbox undo), gt_labels
        [n_gt] (class ids), gt_offsets [bs+1]: image b owns gt rows gt_offsets[b]:gt_offsets[b+1].  Device tensors; one
        host read (the detection total, to size the result)."""
        _lib.require_cuda(rows, "rows")
        dev = rows.device
        bs = offsets.numel() - 1
        total = int(offsets[-1])
        gt_boxes = gt_boxes.to(device=dev, dtype=torch.float32).contiguous()
        gt_labels = gt_labels.to(device=dev, dtype=torch.int32).contiguous()
        gt_offsets = gt_offsets.to(device=dev, dtype=torch.int32).contiguous()
        self._n_gt += torch.bincount(gt_labels.long(), minlength=self.nc)[:self.nc]
        if total == 0:
            return
        rows = rows[:total].contiguous()
        off = offsets.to(device=dev, dtype=torch.int32).contiguous()
        tp = torch.zeros((self.thrs.numel(), total), dtype=torch.uint8, device=dev)
        if gt_boxes.numel() == 0:
            gt_boxes = torch.zeros((1, 4), dtype=torch.float32, device=dev)
            gt_labels = torch.full((1,), -1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.yc_match_detections(rows.data_ptr(), off.data_ptr(), bs, total, gt_boxes.data_ptr(),
                                                    gt_labels.data_ptr(), gt_offsets.data_ptr(), self.thrs.data_ptr(),
                                                    self.thrs.numel(), tp.data_ptr(), _lib.stream_ptr(dev)),
                       "yc_match_detections")
        self._score.append(rows[:, 4] * rows[:, 5])
        self._cls.append(rows[:, 6].long())
        self._tp.append(tp)

    def compute(self):
        """-> dict: 'ap' [n_thr, nc] (NaN for classes without ground truth), 'map' [n_thr], 'map_50_95' (mean over the
        thresholds), all on the device."""
        T = self.thrs.numel()
        ap = torch.full((T, self.nc), float("nan"), dtype=torch.float64, device=self.device)
        has_gt = self._n_gt > 0
        ap[:, has_gt] = 0.0
        if self._score:
            score, cls, tp = torch.cat(self._score), torch.cat(self._cls), torch.cat(self._tp, 1)
            # one stable sort by (class asc, score desc) for all classes at once
            order = torch.sort(score, descending=True, stable=True).indices
            order = order[torch.sort(cls[order], stable=True).indices]
            cls_s, tp_s = cls[order], tp[:, order].to(torch.float64)
            starts = torch.searchsorted(cls_s, torch.arange(self.nc + 1, device=self.device))
            ctp = torch.cumsum(tp_s, 1)
            rec_thr = torch.linspace(0, 1, 101, dtype=torch.float64, device=self.device)
            for c in torch.nonzero(has_gt)[:, 0].tolist():
                lo, hi = int(starts[c]), int(starts[c + 1])
                if hi == lo:
                    continue
                base = ctp[:, lo - 1:lo] if lo > 0 else 0.0
                tpc = ctp[:, lo:hi] - base
                n = torch.arange(1, hi - lo + 1, dtype=torch.float64, device=self.device)
                recall = tpc / float(self._n_gt[c])
                prec = tpc / n
                env = torch.flip(torch.cummax(torch.flip(prec, [1]), 1).values, [1])      # precision envelope
                idx = torch.searchsorted(recall.contiguous(), rec_thr.expand(T, -1).contiguous(), side="left")
                valid = idx < (hi - lo)
                q = torch.gather(env, 1, idx.clamp(max=hi - lo - 1)) * valid
                ap[:, c] = q.mean(1)
        m = torch.nanmean(ap, 1)
        return {"ap": ap, "map": m, "map_50_95": m.mean(), "n_gt": self._n_gt.clone()}
