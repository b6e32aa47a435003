#!/bin/bash
# smoke + default bench line + ncu launch list + one full capture of the head kernel (same command, after it exited 0)
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err; tail -c 3000 gpurun_out/bench_r01g.json
CMD="python bench.py --steps 3 --warmup 3 --profile"
$CMD > gpurun_out/plain_g.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01g.csv $CMD > gpurun_out/ncu_g1.log 2>&1
$CMD > gpurun_out/plain_g2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:head_tc -s 3 -c 2 -o gpurun_out/prof_head_r01g $CMD > gpurun_out/ncu_g2.log 2>&1
tail -n 3 gpurun_out/ncu_g1.log; tail -n 3 gpurun_out/ncu_g2.log
