#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
  timeout 120 python bench.py --steps 200 --warmup 5 --profile 2>&1 | tail -1
done
for d in 5 1; do echo -n "debug=$d "; YC_TC_DEBUG=$d timeout 120 python bench.py --steps 50 --warmup 5 --profile --no-overlap 2>&1 | tail -1; done
