#!/bin/bash
# sweep of tools/mma_probe configurations; output to gpurun_out/probe.log
P=tools/mma_probe
run() { timeout 60 $P "$@" || echo "FAILED: $@"; }
{
for tma in 0 1 2; do
  run 0 256 1 4 $tma 4096
  run 0 256 0 4 $tma 4096
done
run 0 256 1 0 0 4096
run 0 256 1 1 0 4096
run 0 256 1 8 0 4096
run 0 128 1 4 0 4096
run 0 128 1 4 2 4096
run 0 240 1 4 0 4096
run 0 96 1 4 0 4096
for tma in 0 2; do
  run 1 256 1 0 $tma 4096
  run 1 256 1 8 $tma 4096
  run 1 256 1 4 $tma 4096
  run 1 256 1 2 $tma 4096
  run 1 256 0 4 $tma 4096
done
} > gpurun_out/probe.log 2>&1
cat gpurun_out/probe.log
