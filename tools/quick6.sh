#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "fused" 2>&1 | tail -2
YC_TC_2CTA=1 timeout 600 python -m pytest tests -m gpu -x -q -k "fused" 2>&1 | tail -2
for pair in 0 1; do
  echo -n "pair=$pair "
  YC_TC_2CTA=$pair timeout 120 python bench.py --steps 50 --warmup 5 --profile 2>&1 | tail -1
done
