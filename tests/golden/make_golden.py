"""Generate tests/golden/*.npz by importing the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle_golden.py) and are also compared
directly with the CUDA path (tests/test_gpu_*.py).  Harness caveats papered over here
without editing the reference: `stride` is set by hand,
inputs are cloned (both stages mutate them), keep indices are captured by wrapping
the module-global `detect.nms`.
"""
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("YC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")
torch.set_grad_enabled(False)

import detect as ref_detect  # noqa: E402
from nets.detect import Detect  # noqa: E402
from nets.iaux_detect import IAuxDetect  # noqa: E402
from nets.ibin import IBin  # noqa: E402
from nets.idetect import IDetect  # noqa: E402
from utils import bbox as ref_bbox  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
COCO_ANCHORS = [[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]]
TINY_ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
STRIDES = [8.0, 16.0, 32.0]


def trained_like(head, gen, obj_col, cls_from):
    """Weights as SURVEY 8(d): conv N(0,.02) (nets/yolo.py:120), im ~ N(1,.02), bias shift."""
    for name, p in head.named_parameters():
        if name.endswith("weight"):
            p.copy_(torch.randn(p.shape, generator=gen) * 0.02)
        elif name.startswith("im."):
            p.copy_(1.0 + torch.randn(p.shape, generator=gen) * 0.02)
        elif name.startswith("ia."):
            p.copy_(torch.randn(p.shape, generator=gen) * 0.02)
    for conv in head.m:
        b = conv.bias.view(head.na, -1)
        b[:, obj_col] -= 2.0
        b[:, cls_from:] -= 1.0


def sd_np(head, prefix):
    return {prefix + k.replace(".", "__"): v.detach().numpy().copy() for k, v in head.state_dict().items()}


def head_case(cls, name, nc, anchors, ch, shapes, bs, seed, **kw):
    torch.manual_seed(seed)   # default-initialised parameters (conv biases) come from the global generator
    gen = torch.Generator().manual_seed(seed)
    head = cls(nc, anchors, ch, **kw).eval()
    head.stride = torch.tensor(STRIDES)
    if cls is IBin:
        trained_like(head, gen, obj_col=46, cls_from=47)
    else:
        trained_like(head, gen, obj_col=4, cls_from=5)
    xs = [torch.randn(bs, c, h, w, generator=gen) for c, (h, w) in zip(ch, shapes)]
    out = {f"x{i}": x.numpy().copy() for i, x in enumerate(xs)}
    lst = [x.clone() for x in xs]
    z, raw = head(lst)
    out["z"] = z.numpy()
    for i, r in enumerate(raw):
        out[f"raw{i}"] = r.numpy()
    if cls is IAuxDetect:  # the aux maps stay in the caller's list (nets/iaux_detect.py:37-38)
        for i in range(head.nl):
            out[f"aux{i}"] = lst[i + head.nl].numpy()
    out.update(sd_np(head, "sd__"))
    out["nc"] = np.int64(nc)
    out["anchors_cfg"] = np.asarray(anchors, np.float32)
    out["strides"] = np.asarray(STRIDES, np.float32)
    # train-mode output (list of raw maps) must equal the eval raws
    head.train()
    tr = head([x.clone() for x in xs])
    for i in range(head.nl):
        assert torch.equal(tr[i], raw[i])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "z", tuple(z.shape))


def clustered_prediction(bs, rows, nc, seed, n_obj=12, normalised=True):
    """xywh + obj + cls rows with overlapping clusters so that NMS suppresses."""
    g = np.random.default_rng(seed)
    pred = np.zeros((bs, rows, 5 + nc), np.float32)
    for b in range(bs):
        ctr = g.uniform(0.15, 0.85, (n_obj, 2))
        size = g.uniform(0.05, 0.3, (n_obj, 2))
        cls = g.integers(0, nc, n_obj)
        which = g.integers(0, n_obj, rows)
        fg = g.uniform(size=rows) < 0.55
        xy = np.where(fg[:, None], ctr[which] + g.normal(0, 0.012, (rows, 2)) * 1.0, g.uniform(0, 1, (rows, 2)))
        wh = np.where(fg[:, None], size[which] * g.uniform(0.8, 1.2, (rows, 2)), g.uniform(0.01, 0.2, (rows, 2)))
        obj = 1 / (1 + np.exp(-np.where(fg, g.normal(1.5, 1.0, rows), g.normal(-5, 1.0, rows))))
        c = 1 / (1 + np.exp(-g.normal(-3.0, 1.0, (rows, nc))))
        hot = np.where(g.uniform(size=rows) < 0.9, cls[which], g.integers(0, nc, rows))
        c[np.arange(rows), hot] = 1 / (1 + np.exp(-g.normal(2.0, 1.0, rows)))
        pred[b, :, 0:2], pred[b, :, 2:4], pred[b, :, 4], pred[b, :, 5:] = xy, wh, obj, c
    # exact score ties and exactly duplicated boxes, to pin the stable ordering
    pred[0, 7] = pred[0, 3]
    pred[0, 11, 4:] = pred[0, 3, 4:]
    if not normalised:
        pred[..., :4] *= 640.0
    return pred


def nms_case(name, pred, nc, conf, iou, input_shape, image_shape, letterbox):
    records = []
    orig = ref_detect.nms

    def spy(boxes, scores, thr):
        keep = orig(boxes, scores, thr)
        records.append((boxes.numpy().copy(), keep.numpy().copy()))
        return keep

    ref_detect.nms = spy
    try:
        t = torch.from_numpy(pred.copy())
        out = ref_detect.non_max_suppression(t, nc, input_shape, np.array(image_shape), letterbox,
                                             conf_thres=conf, nms_thres=iou)
    finally:
        ref_detect.nms = orig
    after = t.numpy()
    # recover the original row index of every kept detection from the spy records
    keep_idx, counts, rows = [], [], []
    rec = iter(records)
    for b in range(pred.shape[0]):
        img = torch.from_numpy(after[b])
        cc, cp = torch.max(img[:, 5:5 + nc], 1)
        mask = (img[:, 4] * cc >= conf)
        cand = torch.nonzero(mask)[:, 0].numpy()
        idx_b = []
        for c in np.unique(cp[mask].numpy()):
            sub = cand[cp[mask].numpy() == c]
            boxes, keep = next(rec)
            assert boxes.shape[0] == len(sub)
            idx_b.extend(sub[keep].tolist())
        keep_idx.extend(idx_b)
        counts.append(len(idx_b))
        if out[b] is None:
            assert len(idx_b) == 0
        else:
            assert out[b].shape[0] == len(idx_b)
            rows.append(out[b])
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), pred=pred, corners=after[..., :4].copy(),
        rows=(np.concatenate(rows, 0) if rows else np.zeros((0, 7), np.float32)),
        counts=np.asarray(counts, np.int64), keep_idx=np.asarray(keep_idx, np.int64),
        nc=np.int64(nc), conf=np.float64(conf), iou=np.float64(iou),
        input_shape=np.asarray(input_shape, np.int64), image_shape=np.asarray(image_shape, np.int64),
        letterbox=np.int64(1 if letterbox else 0))
    print(name, "counts", counts)


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    # --- a12: the reference's only known-answer data, utils/bbox.py:207-225 ---------------
    xxyy = torch.asarray([[1, 2, 3, 5]]).float()
    kat = {"xxyy": xxyy.numpy()}
    more = torch.tensor([[1, 2, 3, 5], [0.25, 0.5, 0.75, 1.5], [10, 20, 30, 45], [-1, 1, -2, 2]]).float()
    kat["boxes"] = more.numpy()
    for f in ref_bbox.CvtFlag:
        kat[f"kat_{f.value}"] = ref_bbox.cvt_bbox(xxyy, f).numpy()
        kat[f"out_{f.value}"] = ref_bbox.cvt_bbox(more, f).numpy()
    g = torch.Generator().manual_seed(5)
    b1 = torch.rand(17, 4, generator=g); b1[:, 2:] += b1[:, :2]
    b2 = torch.rand(9, 4, generator=g); b2[:, 2:] += b2[:, :2]
    kat["iou_b1"], kat["iou_b2"] = b1.numpy(), b2.numpy()
    kat["iou"] = ref_bbox.box_iou(b1, b2).numpy()
    kat["grid_5_3"] = ref_bbox.make_grid(5, 3).numpy()
    np.savez_compressed(os.path.join(HERE, "bbox_kat.npz"), **kat)

    # --- a10: torchvision.ops.nms semantic probes ------------------------------------------
    from torchvision.ops import nms
    probes = {}
    cases = {
        "third": (torch.tensor([[0, 0, 1, 1], [0, 0, 1, 3]]).float(), torch.tensor([0.9, 0.8])),
        "ties": (torch.tensor([[0, 0, 1, 1], [5, 5, 6, 6], [0, 0, 1, 1.01], [5, 5, 6, 6]]).float(),
                 torch.tensor([0.5, 0.5, 0.5, 0.5])),
        "zero_area": (torch.tensor([[1, 1, 1, 1], [1, 1, 1, 1], [0, 0, 2, 2]]).float(),
                      torch.tensor([0.3, 0.2, 0.1])),
    }
    for k, (bx, sc) in cases.items():
        probes[k + "_boxes"], probes[k + "_scores"] = bx.numpy(), sc.numpy()
        for tn, thr in (("a", 1 / 3), ("b", float(np.float32(1 / 3))), ("c", 0.5)):
            probes[f"{k}_keep_{tn}"] = nms(bx, sc, thr).numpy()
    gg = torch.Generator().manual_seed(11)
    bx = torch.rand(400, 4, generator=gg) * 0.5; bx[:, 2:] = bx[:, :2] + 0.05 + bx[:, 2:] * 0.4
    sc = torch.rand(400, generator=gg); sc[50:60] = sc[40]
    probes["rand_boxes"], probes["rand_scores"] = bx.numpy(), sc.numpy()
    for thr in (0.3, 0.45, 0.65):
        probes[f"rand_keep_{thr}"] = nms(bx, sc, thr).numpy()
    np.savez_compressed(os.path.join(HERE, "nms_probes.npz"), **probes)

    # --- a1-a7: heads ------------------------------------------------------------------------
    head_case(IDetect, "idetect_nc80", 80, COCO_ANCHORS, (16, 32, 64), [(8, 8), (4, 4), (2, 2)], 2, 100)
    head_case(IDetect, "idetect_nc1_rect", 1, TINY_ANCHORS, (8, 16, 32), [(6, 10), (3, 5), (2, 3)], 3, 101)
    head_case(IAuxDetect, "iaux_nc80", 80, COCO_ANCHORS, (16, 32, 64, 16, 32, 64),
              [(8, 8), (4, 4), (2, 2), (8, 8), (4, 4), (2, 2)], 2, 102)
    head_case(IBin, "ibin_nc80", 80, COCO_ANCHORS, (16, 32, 64), [(8, 8), (4, 4), (2, 2)], 2, 103)

    # --- a8: Variant A (Detect conv -> decode_box), detect.py:229 ---------------------------
    gen = torch.Generator().manual_seed(104)
    nc = 3
    det = Detect(nc, TINY_ANCHORS, (8, 16, 32)).eval()
    xs = [torch.randn(2, c, h, w, generator=gen) for c, (h, w) in zip((8, 16, 32), [(8, 8), (4, 4), (2, 2)])]
    for p in det.parameters():
        p.copy_(torch.randn(p.shape, generator=gen) * 0.3)
    convs = det([x.clone() for x in xs])
    anchors = np.asarray(TINY_ANCHORS).reshape(-1, 2)
    mask = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]
    outs = ref_detect.decode_box([c.clone() for c in convs], anchors, mask, nc, image_size=(64, 64))
    va = {f"conv{i}": c.numpy() for i, c in enumerate(convs)}
    va.update({f"out{i}": o.numpy() for i, o in enumerate(outs)})
    va.update({f"x{i}": x.numpy() for i, x in enumerate(xs)})
    va.update(sd_np(det, "sd__"))
    va["anchors"], va["mask"], va["nc"], va["image_size"] = anchors, np.asarray(mask), np.int64(nc), np.asarray([64, 64])
    np.savez_compressed(os.path.join(HERE, "variant_a.npz"), **va)

    # --- 8f: prepare_test_image (detect.py:16-26) on synthetic images (cv2.imread replaced by an in-memory image) ---
    from image_enhance.letter_box import LetterBox
    rng = np.random.default_rng(21)
    for name, (h, w), target in (("letterbox_wide", (48, 77), (64, 64)), ("letterbox_tall_up", (23, 14), (64, 64)),
                                 ("letterbox_same", (64, 64), (64, 64)), ("letterbox_down", (300, 171), (96, 96))):
        # smooth-ish content plus noise so that interpolation errors would show
        yy, xx = np.mgrid[0:h, 0:w]
        base = (127 + 100 * np.sin(xx / 5.0)[..., None] * np.cos(yy / 7.0)[..., None] * np.ones(3)).astype(np.float64)
        img = np.clip(base + rng.integers(-30, 31, (h, w, 3)), 0, 255).astype(np.uint8)
        image_data, _ = LetterBox(target, scale_fill_prob=0)(img, np.zeros((0, 4)))
        data = np.expand_dims(np.transpose((np.array(image_data, dtype='float32') / 255.), (2, 0, 1)), 0)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), image=img, target=np.asarray(target), data=data)

    # --- 8f rank 4: RepConv.fuse_repvgg_block (nets/common.py:565-614) with and without the identity branch ---------
    import copy
    from nets.common import RepConv
    for name, (c1, c2, s_) in (("repconv_identity", (8, 8, 1)), ("repconv_noid", (6, 10, 2))):
        torch.manual_seed(31)   # RepConv's conv weights keep their default initialisation (global generator)
        gen = torch.Generator().manual_seed(31)
        rep = RepConv(c1, c2, 3, s_).eval()
        for m_ in rep.modules():
            if isinstance(m_, torch.nn.BatchNorm2d):
                m_.running_mean.copy_(torch.randn(m_.running_mean.shape, generator=gen) * 0.3)
                m_.running_var.copy_(torch.rand(m_.running_var.shape, generator=gen) + 0.5)
                m_.weight.copy_(1.0 + 0.2 * torch.randn(m_.weight.shape, generator=gen))
                m_.bias.copy_(0.1 * torch.randn(m_.bias.shape, generator=gen))
        x = torch.randn(2, c1, 9, 7, generator=gen)
        before = rep(x).numpy()
        sd = sd_np(rep, "sd__")
        fused = copy.deepcopy(rep)
        fused.fuse_repvgg_block()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), before=before, after=fused(x).numpy(),
                            weight=fused.rbr_reparam.weight.detach().numpy(), bias=fused.rbr_reparam.bias.detach().numpy(),
                            c=np.asarray([c1, c2, s_]), **sd)

    # --- 8f rank 1: the formatting loop of predict (detect.py:236-258).  It lives inside predict(), which cannot run
    # (Windows checkpoint path, cv2.imshow), so the reference's own source lines of that block are executed here on
    # synthetic NMS results, with stand-ins for the names the block reads. ---------------------------------------
    import inspect
    import textwrap
    import types
    src = inspect.getsource(ref_detect.predict)
    block = textwrap.dedent(src[src.index("    if results[0] is not None:"):src.index("        show_bbox(original_image, target_boxes)")])
    rng = np.random.default_rng(41)
    n, (ih, iw) = 40, (480, 640)
    rows = np.concatenate([rng.uniform(-30, 700, (n, 4)), rng.uniform(0, 1, (n, 2)), rng.integers(0, 80, (n, 1))], 1).astype(np.float32)
    rows[0, :4] = [-0.5, -0.0, 479.999, 640.0]          # edge values of floor / clamp
    rows[1, :4] = [3.0, 5.0, 480.0, 639.5]
    ns = {"np": np, "results": [rows.copy()], "original_image": np.zeros((ih, iw, 3), np.uint8),
          "plan": types.SimpleNamespace(labels=[str(i) for i in range(80)]), "colors": [(0, 0, 0)] * 80,
          "TargetBox": ref_detect.TargetBox, "print": lambda *a, **k: None}
    exec(block, ns)
    tb = ns["target_boxes"]
    np.savez_compressed(os.path.join(HERE, "format_predict.npz"), rows=rows, image_hw=np.asarray([ih, iw]),
                        box=np.asarray([[t.left, t.top, t.right, t.bottom] for t in tb], dtype=np.int64),
                        conf=np.asarray([t.score for t in tb], dtype=np.float32),
                        label=np.asarray([int(t.label) for t in tb], dtype=np.int64))

    # --- a9-a11: NMS ---------------------------------------------------------------------------
    nms_case("nms_clustered_lb", clustered_prediction(2, 320, 80, 7), 80, 0.25, 0.45, (640, 640), (512, 773), True)
    nms_case("nms_clustered_nolb", clustered_prediction(2, 320, 80, 8), 80, 0.3, 0.3, (640, 640), (480, 640), False)
    nms_case("nms_lowconf", clustered_prediction(1, 400, 80, 9), 80, 0.001, 0.65, (640, 640), (640, 640), True)
    p = clustered_prediction(3, 64, 4, 10)
    p[1, :, 4] = 0.0  # image 1 has no detection -> None
    nms_case("nms_with_none", p, 4, 0.5, 0.4, (640, 640), (300, 500), True)
    nms_case("nms_nc1", clustered_prediction(2, 200, 1, 12, n_obj=5), 1, 0.3, 0.3, (640, 640), (512, 773), True)
    # threshold rounding: 0.7 is not representable; fp32(0.7) < 0.7 and torch compares in fp32
    p = clustered_prediction(1, 96, 2, 13)
    p[0, 0, 4], p[0, 0, 5:] = np.float32(0.7), np.float32([1.0, 0.1])
    nms_case("nms_thr_round", p, 2, 0.7, 0.5, (640, 640), (640, 640), True)


if __name__ == "__main__":
    main()
