#!/bin/bash
# quick GPU check: tensor-core head parity, then head timing for the 1-CTA and the CTA-pair kernel
timeout 600 python -m pytest tests -m gpu -x -q -k "tcgen05 or fused" 2>&1 | tail -3
YC_TC_2CTA=1 timeout 600 python -m pytest tests -m gpu -x -q -k "fused" 2>&1 | tail -3
for pair in 0 1; do
for d in 0 1 3 5 7; do
  echo -n "pair=$pair debug=$d "
  YC_TC_2CTA=$pair YC_TC_DEBUG=$d timeout 120 python bench.py --steps 50 --warmup 5 --profile 2>&1 | tail -1
done
done
