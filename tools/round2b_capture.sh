#!/bin/bash
# round 2, second session: smoke + default bench line + ncu launch list + full captures of the head kernels that changed
# (each ncu run only after the same command exited 0 without ncu); outputs under gpurun_out/, summarised into profiles/
# with profiles/summarize.py
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; tail -2 gpurun_out/r02b_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 300 gpurun_out/r02b_bench.err
CMD="python bench.py --steps 3 --warmup 3 --profile"
$CMD > gpurun_out/r02b_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_S3.csv $CMD > gpurun_out/r02b_ncu1.log 2>&1
$CMD > gpurun_out/r02b_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:head_tc2_kernel -s 3 -c 1 -f -o gpurun_out/r02b_prof_head_tc2_fused $CMD > gpurun_out/r02b_ncu2.log 2>&1
export CASES=s1
CMD2="python tools/fwd_time.py"
$CMD2 > gpurun_out/r02b_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:head_tc2_kernel -s 5 -c 1 -f -o gpurun_out/r02b_prof_head_tc2_z $CMD2 > gpurun_out/r02b_ncu3.log 2>&1
export CASES=ibin
$CMD2 > gpurun_out/r02b_plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:head_tc2i_kernel -s 5 -c 1 -f -o gpurun_out/r02b_prof_head_tc2i_z $CMD2 > gpurun_out/r02b_ncu4.log 2>&1
CMD3="python tools/c5_time.py"
$CMD3 > gpurun_out/r02b_plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:head_tc2i_kernel -s 3 -c 1 -f -o gpurun_out/r02b_prof_head_tc2i_fused $CMD3 > gpurun_out/r02b_ncu5.log 2>&1
tail -n 2 gpurun_out/r02b_ncu1.log gpurun_out/r02b_ncu2.log gpurun_out/r02b_ncu3.log gpurun_out/r02b_ncu4.log gpurun_out/r02b_ncu5.log
