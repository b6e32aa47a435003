#!/bin/bash
YC_TC_2CTA=1 timeout 300 python -m pytest tests -m gpu -x -q -k "fused" 2>&1 | tail -30
echo ==== sanitizer
YC_TC_2CTA=1 timeout 600 compute-sanitizer --tool memcheck python bench.py --steps 1 --warmup 3 --profile --bs 4 2>&1 | grep -v "^$" | head -60
