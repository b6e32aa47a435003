"""Debug probe of the float32 split kernel: one forward on the tcgen05 path against plain torch float64 math."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_continuous_b200 import _lib
from yolo_continuous_b200.nets import IDetect

dev = "cuda:0"
COCO = [[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]]
ch = tuple(int(c) for c in os.environ.get("CH", "256,512,1024").split(","))
hw = int(os.environ.get("HW", "16"))
bs = int(os.environ.get("BS", "2"))
g = torch.Generator().manual_seed(0)
head = IDetect(80, COCO, ch).eval()
with torch.no_grad():
    for n, p in head.named_parameters():
        if n.endswith("weight"):
            p.copy_(0.02 * torch.randn(p.shape, generator=g))
        elif n.startswith("im."):
            p.copy_(1.0 + 0.02 * torch.randn(p.shape, generator=g))
        elif n.startswith("ia."):
            p.copy_(0.02 * torch.randn(p.shape, generator=g))
head.stride = torch.tensor([8.0, 16.0, 32.0])
head = head.to(dev)
xs = [torch.randn(bs, c, hw, hw, generator=g).to(dev) for c in ch]
head.head_path = _lib.YC_PATH_TCGEN05
_, raws = head(list(xs))
torch.cuda.synchronize()
for i in range(3):
    w = head.m[i].weight.double()[:, :, 0, 0]
    x = xs[i].double() + head.ia[i].implicit.double()
    ref = torch.einsum("nk,bkhw->bnhw", w, x) + head.m[i].bias.double().view(1, -1, 1, 1)
    ref = ref * head.im[i].implicit.double()
    ref = ref.view(bs, 3, 85, hw, hw).permute(0, 1, 3, 4, 2)
    err = (raws[i].double() - ref).abs()
    print(f"level {i} K={ch[i]}: max |err| {err.max().item():.3e}  mean {err.mean().item():.3e}  max |ref| {ref.abs().max().item():.2f}")
