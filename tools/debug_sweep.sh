#!/bin/bash
# head kernel time decomposition: YC_TC_DEBUG bits 1 = skip epilogue, 2 = skip MMA issue, 4 = skip TMA
for pair in 0 1; do
for d in 0 1 2 3 4 5 6 7; do
  echo -n "pair=$pair debug=$d "
  YC_TC_2CTA=$pair YC_TC_DEBUG=$d timeout 120 python bench.py --steps 50 --warmup 5 --profile 2>&1 | tail -1
done
done > gpurun_out/debug_sweep.log 2>&1
cat gpurun_out/debug_sweep.log
