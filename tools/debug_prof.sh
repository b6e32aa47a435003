#!/bin/bash
for d in 15 11 13 9 8; do
  echo "== debug=$d"
  YC_TC_2CTA=0 YC_TC_DEBUG=$d timeout 120 python bench.py --steps 2 --warmup 3 --profile 2>&1 | tail -3
done > gpurun_out/debug_prof.log 2>&1
cat gpurun_out/debug_prof.log
