"""Pin the CPU oracle (oracle/) against fixtures produced by the unmodified reference.

Bit-exact where the arithmetic is IEEE-determined (layout conversions, corners,
thresholding, NMS keep sets, letterbox undo); tolerance 2e-6 (scaled, see helpers)
where libm/oneDNN differ by an ulp or two (1x1 conv accumulation order, expf).
"""
import numpy as np
import pytest

from oracle import oracle as orc
from helpers import assert_close_scaled, box_scale, head_params, load

RTOL_ORACLE = 2e-6


@pytest.mark.parametrize("name,kind", [("idetect_nc80", "idetect"), ("idetect_nc1_rect", "idetect"),
                                       ("iaux_nc80", "iaux"), ("ibin_nc80", "ibin")])
def test_head_forward_matches_reference(name, kind):
    fx = load(name)
    p = head_params(fx, kind)
    nl = p["anchors"].shape[0]
    nx_in = 2 * nl if kind == "iaux" else nl
    xs = [fx[f"x{i}"] for i in range(nx_in)]
    res = orc.head_forward(kind, p, xs, fx["strides"])
    z, raws = res[0], res[1]
    shapes = [fx[f"raw{i}"].shape[2:4] for i in range(nl)]
    na = p["anchors"].shape[1]
    for i in range(nl):
        np.testing.assert_allclose(raws[i], fx[f"raw{i}"], rtol=0, atol=RTOL_ORACLE, err_msg=f"raw{i}")
    assert z.shape == fx["z"].shape
    scale = box_scale(p["anchors"], fx["strides"], shapes, na, z.shape[-1])
    assert_close_scaled(z, fx["z"], scale, RTOL_ORACLE, name)
    if kind == "iaux":
        for i in range(nl):
            np.testing.assert_allclose(res[2][i], fx[f"aux{i}"], rtol=0, atol=RTOL_ORACLE)


def test_decode_box_variant_a():
    fx = load("variant_a")
    outs = orc.decode_box([fx[f"conv{i}"] for i in range(3)], fx["anchors"], fx["mask"].tolist(),
                          int(fx["nc"]), tuple(fx["image_size"]))
    for i in range(3):
        np.testing.assert_allclose(outs[i], fx[f"out{i}"], rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize("name", ["nms_clustered_lb", "nms_clustered_nolb", "nms_lowconf",
                                  "nms_with_none", "nms_nc1", "nms_thr_round"])
def test_non_max_suppression_bit_exact(name):
    fx = load(name)
    pred = fx["pred"].copy()
    out, idx = orc.non_max_suppression(pred, int(fx["nc"]), tuple(fx["input_shape"]), tuple(fx["image_shape"]),
                                       bool(fx["letterbox"]), float(fx["conf"]), float(fx["iou"]),
                                       return_indices=True)
    # in-place corner write-back, detect.py:98-103
    assert np.array_equal(pred[..., :4], fx["corners"])
    counts = [0 if o is None else o.shape[0] for o in out]
    assert counts == fx["counts"].tolist()
    assert np.array_equal(np.concatenate(idx), fx["keep_idx"])
    rows = [o for o in out if o is not None]
    got = np.concatenate(rows, 0) if rows else np.zeros((0, 7), np.float32)
    assert np.array_equal(got[:, 4:], fx["rows"][:, 4:])
    assert np.array_equal(got[:, :4], fx["rows"][:, :4].astype(np.float32))
    for o, c in zip(out, counts):
        assert (o is None) == (c == 0)


def test_nms_probes():
    fx = load("nms_probes")
    for k in ("third", "ties", "zero_area"):
        for tn, thr in (("a", 1 / 3), ("b", float(np.float32(1 / 3))), ("c", 0.5)):
            keep = orc.nms(fx[k + "_boxes"], fx[k + "_scores"], thr)
            assert np.array_equal(keep, fx[f"{k}_keep_{tn}"]), (k, tn)
    for thr in (0.3, 0.45, 0.65):
        assert np.array_equal(orc.nms(fx["rand_boxes"], fx["rand_scores"], thr), fx[f"rand_keep_{thr}"])


def test_bbox_known_answers():
    """utils/bbox.py:207-225, the reference's only known-answer data."""
    fx = load("bbox_kat")
    expect = {0: [1, 3, 2, 5], 1: [1.5, 4, 1, 2]}
    assert orc.cvt_bbox(fx["xxyy"], 0).tolist() == [expect[0]]
    assert orc.cvt_bbox(fx["xxyy"], 1).tolist() == [expect[1]]
    for f in range(6):
        assert np.array_equal(orc.cvt_bbox(fx["boxes"], f), fx[f"out_{f}"]), f
    with pytest.raises(Exception):
        orc.cvt_bbox(fx["boxes"], 6)
    assert np.array_equal(orc.box_iou(fx["iou_b1"], fx["iou_b2"]), fx["iou"])


def test_torch_port_matches_reference_bit_exact():
    """oracle/ref_port.py (the CPU-baseline port) reproduces the reference's outputs exactly."""
    import torch
    from oracle import ref_port
    fx = load("idetect_nc80")
    p = head_params(fx, "idetect")
    z, raws = ref_port.idetect_forward(p, [fx[f"x{i}"] for i in range(3)], fx["strides"])
    assert np.array_equal(z.numpy(), fx["z"])
    for i in range(3):
        assert np.array_equal(raws[i].numpy(), fx[f"raw{i}"])
    for name in ("nms_clustered_lb", "nms_clustered_nolb", "nms_with_none", "nms_thr_round"):
        fx = load(name)
        out = ref_port.non_max_suppression(torch.from_numpy(fx["pred"].copy()), int(fx["nc"]), tuple(fx["input_shape"]),
                                           np.array(fx["image_shape"]), bool(fx["letterbox"]), float(fx["conf"]),
                                           float(fx["iou"]))
        rows = [o for o in out if o is not None]
        got = np.concatenate(rows, 0) if rows else np.zeros((0, 7), np.float32)
        assert np.array_equal(got, fx["rows"].astype(np.float32))
        assert [0 if o is None else len(o) for o in out] == fx["counts"].tolist()


LETTERBOX_FIXTURES = ["letterbox_wide", "letterbox_tall_up", "letterbox_same", "letterbox_down"]


@pytest.mark.parametrize("name", LETTERBOX_FIXTURES)
def test_oracle_prepare_test_image_bit_exact(name):
    """prepare_test_image (detect.py:16-26: LetterBox + cv2.resize INTER_LINEAR + /255 + CHW) restated without
    OpenCV, bit for bit against the reference's own output."""
    fx = load(name)
    got = orc.prepare_test_image(fx["image"], tuple(int(v) for v in fx["target"]))
    assert got.dtype == np.float32 and got.shape == fx["data"].shape
    assert np.array_equal(got, fx["data"])


def test_oracle_format_detections_vs_reference_block():
    """The reference's own formatting lines (detect.py:236-258, executed on synthetic rows by make_golden.py)."""
    fx = load("format_predict")
    box, conf, label = orc.format_detections(fx["rows"], tuple(int(v) for v in fx["image_hw"]))
    assert np.array_equal(box, fx["box"]) and np.array_equal(conf, fx["conf"]) and np.array_equal(label, fx["label"])


def test_oracle_format_detections():
    """Formatting loop of predict (detect.py:236-258): floor, clamp, obj*cls_conf, int label."""
    rows = np.array([[-3.7, 10.2, 50.9, 700.5, 0.9, 0.5, 3.0],
                     [12.0, -0.4, 480.0, 639.99, 0.25, 0.25, 79.0],
                     [470.5, 630.5, 490.0, 650.0, 1.0, 0.3, 0.0]], np.float32)
    box, conf, label = orc.format_detections(rows, (480, 640))
    assert box.tolist() == [[10, 0, 640, 50], [0, 12, 639, 480], [630, 470, 640, 480]]
    assert np.array_equal(conf, rows[:, 4] * rows[:, 5]) and label.tolist() == [3, 79, 0]


@pytest.mark.parametrize("name", ["repconv_identity", "repconv_noid"])
def test_repconv_reparam_vs_reference(name):
    """RepConv.fuse_repvgg_block (nets/common.py:565-614): the collapsed 3x3 weight / bias equal the reference's bit for
    bit and the fused block reproduces its output (host logic: runs wherever the parameters live)."""
    import torch
    from helpers import make_rep
    from yolo_continuous_b200.nets.common import fuse_repvgg_block, repconv_equivalent
    fx = load(name)
    c1, c2, s_ = (int(v) for v in fx["c"])

    rep = make_rep(c1, c2, s_)   # the reference's layout (nets/common.py:440-472), rebuilt from the fixture
    rep.load_state_dict({k[4:].replace("__", "."): torch.from_numpy(v) for k, v in fx.items() if k.startswith("sd__")})
    x = torch.from_numpy(fx["x"])
    with torch.no_grad():
        assert np.array_equal(rep(x).numpy(), fx["before"])
        w, b = repconv_equivalent(rep)
        assert np.array_equal(w.numpy(), fx["weight"]) and np.array_equal(b.numpy(), fx["bias"])
        fuse_repvgg_block(rep)
        assert rep.deploy and rep.rbr_dense is None and rep.rbr_1x1 is None and rep.rbr_identity is None
        assert np.array_equal(rep(x).numpy(), fx["after"])


def test_oracle_nms_matches_torchvision_on_random_boxes():
    """torchvision.ops.nms is the third-party routine behind detect.py:133; where the wheel is installed (it is in
    the build image) the C restatement is checked against it directly on seeded random boxes with exact ties,
    duplicates and degenerate boxes, at thresholds that are and are not representable in binary32."""
    tv = pytest.importorskip("torchvision")
    import torch
    rng = np.random.default_rng(17)
    for n, thr in ((1, 0.5), (7, 0.3), (64, 0.45), (257, 0.65), (500, 0.5), (300, 1.0 / 3.0), (120, 0.0)):
        xy = rng.uniform(0, 100, (n, 2)).astype(np.float32)
        wh = rng.uniform(0, 40, (n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = rng.uniform(0, 1, n).astype(np.float32)
        if n > 10:
            boxes[5] = boxes[3]                      # duplicate box
            scores[5] = scores[3]                    # ... with a tied score (stable order decides)
            boxes[7, 2:] = boxes[7, :2]              # zero-area box
            scores[8] = scores[9]                    # tie between different boxes
            boxes[11] = np.round(boxes[11])          # integer coordinates: IoU exactly representable more often
            boxes[12] = boxes[11] + np.float32([0, 0, 10, 0])
        want = tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
        got = orc.nms(boxes, scores, thr)
        assert np.array_equal(got, want), (n, thr)


def test_oracle_evaluator_hand_computed_case():
    """evaluate_map on a case small enough to do by hand: one class, two ground-truth boxes, three detections
    (hit, duplicate of the same box = false positive, hit): AP@0.5 = mean over recall points of the envelope."""
    gt = (np.array([[0, 0, 10, 10], [20, 20, 30, 30]], np.float32), np.array([0, 0]))
    det = np.array([[0, 0, 10, 10, 1.0, 0.9, 0], [0, 0, 10, 9, 1.0, 0.8, 0], [20, 20, 30, 31, 1.0, 0.7, 0]], np.float32)
    tps, ap = orc.evaluate_map([det], [gt], 2, [0.5, 0.95])
    assert tps[0].tolist() == [[1, 0, 1], [1, 0, 0]]
    # thr 0.5: precision 1, 1/2, 2/3 at recall .5, .5, 1 -> envelope 1 up to recall .5 (51 points), 2/3 beyond (50 points)
    assert abs(ap[0, 0] - (51 * 1.0 + 50 * 2 / 3) / 101) < 1e-12
    assert abs(ap[1, 0] - 51 / 101) < 1e-12 and np.isnan(ap[:, 1]).all()
