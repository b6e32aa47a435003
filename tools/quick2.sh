#!/bin/bash
YC_TC_2CTA=1 timeout 600 python -m pytest tests -m gpu -x -q -k "fused" 2>&1 | tail -2
for pair in 1; do
for d in 0 1 5 3; do
  echo -n "pair=$pair debug=$d "
  YC_TC_2CTA=$pair YC_TC_DEBUG=$d timeout 120 python bench.py --steps 50 --warmup 5 --profile 2>&1 | tail -1
done
done
