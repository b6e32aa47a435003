"""Forward (z-materialising) timings of the drop-in heads, one line per case; used with the YC_TC_DEBUG / YC_TS_DEBUG
switches of the head kernels.  CASES=s1,s1raw,fp32,iaux,ibin selects; BS overrides the batch of the C2 cases."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yolo_continuous_b200.nets import IAuxDetect, IBin

dev = torch.device("cuda:0")
cases = os.environ.get("CASES", "s1,fp32,iaux,ibin").split(",")
tag = f"TC={os.environ.get('YC_TC_DEBUG', '0')} TS={os.environ.get('YC_TS_DEBUG', '0')}"
bs = int(os.environ.get("BS", "64"))


def run(name, head, xs, byts):
    with torch.no_grad():
        ms = bench.timed_gpu(lambda: head(list(xs)), 20, 3)
    print(f"{tag} {name}: {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s algorithmic ({byts / ms / 1e6 / 6548.8:.3f} of HBM)", flush=True)


for c in cases:
    if c in ("s1", "s1raw", "fp32"):
        head = bench.make_head().to(dev)
        head.return_raw = c == "s1raw"
        dt = torch.float32 if c == "fp32" else torch.bfloat16
        xs = bench.make_maps(bs, 1234, dt, dev)
        byts = bench.BYTES_PER_IMG["fp32" if c == "fp32" else "bf16"] * bs + (25200 * 85 * 4 * bs if c == "s1raw" else 0)
        run(c, head, xs, byts)
    else:
        shapes5 = [(160, 160), (80, 80), (40, 40)]
        cls, ch = (IAuxDetect, bench.CH * 2) if c == "iaux" else (IBin, bench.CH)
        head = cls(bench.NC, bench.COCO_ANCHORS, ch).eval().to(dev)
        head.stride = torch.tensor(bench.STRIDES)
        head.return_raw = False
        head.compute_aux_in_eval = False
        xs = bench.make_maps(16, 7, torch.bfloat16, dev, ch, shapes5 * (len(ch) // 3))
        elems5 = sum(k * h * w for k, (h, w) in zip(bench.CH, shapes5))
        byts = (elems5 * 2 + 100800 * 85 * 4) * 16
        run(c, head, xs, byts)
    del head, xs
    torch.cuda.empty_cache()
