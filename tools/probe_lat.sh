#!/bin/bash
P=tools/mma_probe
run() { timeout 60 $P "$@" || echo "FAILED: $@"; }
{
run 0 256 1 0 0 2000 1
run 0 256 1 1 0 2000 1
run 0 256 1 4 0 2000 1
run 0 256 1 8 0 2000 1
run 1 256 1 0 0 2000 1
run 1 256 1 1 0 2000 1
run 1 256 1 4 0 2000 1
run 0 256 1 0 0 2000 2
run 0 256 1 0 0 2000 3
} > gpurun_out/probe_lat.log 2>&1
cat gpurun_out/probe_lat.log
