#!/bin/bash
# decomposition of the z-materialising forwards with the debug switches of the head kernels
for d in 0 1 2 3 4 5; do
  YC_TC_DEBUG=$d CASES=s1,iaux,ibin timeout 200 python tools/fwd_time.py 2>&1 | grep -v Warning
done > gpurun_out/fwd_sweep.log 2>&1
cat gpurun_out/fwd_sweep.log
