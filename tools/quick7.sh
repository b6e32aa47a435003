#!/bin/bash
for rep in 1 2; do for pair in 0 1; do
  echo -n "pair=$pair "
  YC_TC_2CTA=$pair timeout 120 python bench.py --steps 50 --warmup 5 --profile 2>&1 | tail -1
done; done
