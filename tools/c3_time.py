"""C3 (mAP-eval thresholds, conf 0.001 / iou 0.65, batch 256, trained-like parameters): a few fused steps, for the ncu
launch list of the NMS kernels (`ncu --metrics gpu__time_duration.sum ... python tools/c3_time.py`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yolo_continuous_b200.pipeline import PostBackbone

dev = torch.device("cuda:0")
bs = int(os.environ.get("BS", "256"))
head = bench.make_head(bench.make_params()).to(dev)
xs = bench.make_maps(bs, 1234, torch.bfloat16, dev)
pipe = PostBackbone(head, bs, bench.SHAPES, torch.bfloat16, bench.INPUT_SHAPE, bench.IMAGE_SHAPE, True, 0.001, 0.65, dev,
                    use_graph=False, overlap=False)
ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(4)]
for e in ev:
    pipe.run_device(xs, head_events=e)
pipe.wait()
torch.cuda.synchronize()
print(f"C3 bs {bs}: head kernel {sorted(e[0].elapsed_time(e[1]) for e in ev)[2] * 1e3:.0f} us, NMS kernels "
      f"{sorted(e[2].elapsed_time(e[3]) for e in ev)[2] * 1e3:.0f} us, detections {int(pipe.meta[bs:][-1])}")
