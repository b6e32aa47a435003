#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or overlapped" 2>&1 | tail -2
for rep in 1 2; do
  timeout 120 python bench.py --steps 200 --warmup 5 --profile 2>&1 | tail -1
done
