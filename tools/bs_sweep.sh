#!/bin/bash
for pair in 0 1; do
for d in 7 0; do
for bs in 16 32 64 128; do
  echo -n "pair=$pair debug=$d bs=$bs "
  YC_TC_2CTA=$pair YC_TC_DEBUG=$d timeout 120 python bench.py --steps 50 --warmup 5 --profile --bs $bs 2>&1 | tail -1
done; done; done > gpurun_out/bs_sweep.log 2>&1
cat gpurun_out/bs_sweep.log
