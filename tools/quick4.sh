#!/bin/bash
for d in 128 144 160 176 0; do
  echo -n "pair=1 debug=$d "
  YC_TC_2CTA=1 YC_TC_DEBUG=$d timeout 120 python bench.py --steps 50 --warmup 5 --profile 2>&1 | tail -1
done
