#!/bin/bash
# full GPU check: all parity tests with the default kernels, fused tests again with the CTA-pair kernel, one bench line
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
YC_TC_2CTA=1 timeout 600 python -m pytest tests -m gpu -x -q -k "fused" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; tail -c 2500 gpurun_out/bench_cur.json
