#!/bin/bash
# ncu --set full of the stand-alone threshold/compaction kernel (S2) and the three NMS kernels, unfused C2 step
CMD="python bench.py --steps 2 --warmup 3 --profile --unfused"
$CMD > gpurun_out/plain_t.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"threshold_compact|bucket_kernel|nms_segment|finish_kernel" -s 12 -c 4 -o gpurun_out/prof_tail_r01 $CMD > gpurun_out/ncu_t.log 2>&1
tail -2 gpurun_out/ncu_t.log
