"""Worker of tests/test_integration_cpu.py (own process: it imports the reference checkout, whose top-level module names
-- nets, utils, detect -- must not leak into the test session).  INTEGRATION.md section 1, exercised: the reference's
model builder (nets/yolo.py:15-87 parse_model, :95-112 Model) with the head names rebound to the B200 drop-ins."""
import os
import sys

import torch
import yaml

REF = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import nets.yolo as ref_yolo  # noqa: E402  (the reference's builder)
from yolo_continuous_b200 import nets as b200  # noqa: E402

ANCHORS = [[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]]
cfg = yaml.safe_load(open(os.path.join(REF, "cfg", "net", "yolov7-tiny.yaml")))
assert cfg["head"][-1][2] == "Detect"
ref_classes = {n: getattr(ref_yolo, n) for n in ("IDetect", "IAuxDetect", "IBin", "Detect")}

for name, nc in (("IDetect", 1), ("IDetect", 80), ("IBin", 80), ("Detect", 3)):
    c = {k: (list(map(list, v)) if k in ("backbone", "head") else v) for k, v in cfg.items()}
    c["head"][-1] = [c["head"][-1][0], 1, name, ["nc", "anchors"]]   # the YAML seam: the head row names the class
    torch.manual_seed(0)
    ref_model = ref_yolo.Model(c, ANCHORS, nc)                       # the reference's own head class
    sd = ref_model.state_dict()
    try:
        for n in ref_classes:                                        # rebind the names
            setattr(ref_yolo, n, getattr(b200, n))
        torch.manual_seed(0)
        model = ref_yolo.Model(c, ANCHORS, nc)
    finally:
        for n, cls in ref_classes.items():
            setattr(ref_yolo, n, cls)
    head, ref_head = model.model[-1], ref_model.model[-1]
    assert type(head) is getattr(b200, name) and type(ref_head) is ref_classes[name]
    # what parse_model tags every module with (nets/yolo.py:81) and what it derived for the head
    assert (head.i, head.f, head.np) == (ref_head.i, ref_head.f, ref_head.np)
    assert head.type.endswith(name)
    if name != "Detect":
        assert (head.nc, head.no, head.nl, head.na) == (ref_head.nc, ref_head.no, ref_head.nl, ref_head.na)
        assert [m.weight.shape for m in head.m] == [m.weight.shape for m in ref_head.m]
        assert head.stride is None and ref_head.stride is None      # nobody sets it (nets/idetect.py:8)
    # checkpoints: identical keys and shapes, strict load
    assert list(model.state_dict().keys()) == list(sd.keys())
    assert all(model.state_dict()[k].shape == v.shape for k, v in sd.items())
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert all(torch.equal(model.state_dict()[k], v) for k, v in sd.items())
    # Model.initial_weights (nets/yolo.py:114-125) re-initialises the head convolutions of both alike
    print("ok", name, nc, "params", head.np)
# the eval forward of the rebound model needs a GPU (no CPU fallback): make sure it says so instead of computing
try:
    model.eval()
    for n in ("IDetect", "IAuxDetect", "IBin"):
        setattr(ref_yolo, n, getattr(b200, n))
    c = {k: (list(map(list, v)) if k in ("backbone", "head") else v) for k, v in cfg.items()}
    c["head"][-1] = [c["head"][-1][0], 1, "IDetect", ["nc", "anchors"]]
    m2 = ref_yolo.Model(c, ANCHORS, 1).eval()
    m2.model[-1].stride = torch.tensor([8., 16., 32.])
    try:
        m2(torch.zeros(1, 3, 64, 64))
        raise SystemExit("CPU forward did not raise")
    except Exception as e:   # YcError: head input must be a CUDA tensor
        assert "CUDA" in str(e), e
finally:
    for n, cls in ref_classes.items():
        setattr(ref_yolo, n, cls)
print("INTEGRATION_OK")
