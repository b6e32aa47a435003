#!/bin/bash
for conf in 0.25 0.9999; do for pair in 0 1; do
  echo -n "conf=$conf pair=$pair "
  YC_BENCH_CONF=$conf YC_TC_2CTA=$pair timeout 120 python bench.py --steps 50 --warmup 5 --profile 2>&1 | tail -1
done; done
