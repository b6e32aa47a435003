#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
  echo -n "overlap "; timeout 120 python bench.py --steps 100 --warmup 5 --profile 2>&1 | tail -1
  echo -n "serial  "; timeout 120 python bench.py --steps 100 --warmup 5 --profile --no-overlap 2>&1 | tail -1
done
