#!/bin/bash
for d in 7 23 39 71 55 119 128 144 176 240; do
  echo -n "debug=$d "
  YC_TC_2CTA=0 YC_TC_DEBUG=$d timeout 120 python bench.py --steps 50 --warmup 5 --profile 2>&1 | tail -1
done > gpurun_out/poll_sweep.log 2>&1
cat gpurun_out/poll_sweep.log
