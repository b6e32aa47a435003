#!/bin/bash
# round 2: smoke + default bench line + ncu launch list + full captures of the head kernels (each ncu run only after the
# same command exited 0 without ncu); outputs under gpurun_out/, summarised into profiles/ with profiles/summarize.py
python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench.err
CMD="python bench.py --steps 3 --warmup 3 --profile"
$CMD > gpurun_out/r02_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_S3.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
$CMD > gpurun_out/r02_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:head_tc2 -s 3 -c 1 -o gpurun_out/r02_prof_head_tc2 $CMD > gpurun_out/r02_ncu2.log 2>&1
CMD2="python tools/split_time.py"
$CMD2 > gpurun_out/r02_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:head_tcs -s 2 -c 1 -o gpurun_out/r02_prof_head_tcs $CMD2 > gpurun_out/r02_ncu3.log 2>&1
tail -n 2 gpurun_out/r02_ncu1.log gpurun_out/r02_ncu2.log gpurun_out/r02_ncu3.log
