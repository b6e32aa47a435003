"""Timing of the float32 split head kernel on the C2 batch (z only) -- used with the YC_TS_DEBUG switches
(1 skip epilogue, 2 skip MMAs, 16 skip feature-map loads, 32 skip weight loads, 64 skip the conversion)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
head = bench.make_head().to(dev)
head.return_raw = False
bs = int(os.environ.get("BS", "64"))
xs = bench.make_maps(bs, 1234, torch.float32, dev)
with torch.no_grad():
    ms = bench.timed_gpu(lambda: head(list(xs)), 20, 3)
print(f"YC_TS_DEBUG={os.environ.get('YC_TS_DEBUG', '0')}: {ms * 1e3:.1f} us per {bs} images = {bs / ms:.1f} k img/s")
