#!/bin/bash
# per-warp cycle breakdown of the z-writing epilogue (YC_TC_DEBUG bit 8) on one forward
export YC_LIB_PATH=${YC_LIB_PATH:-$PWD/exp/lib_dbg.so}
for d in ${DBG:-8}; do
  for c in ${CASES:-s1 ibin}; do
    echo "== debug $d case $c"
    YC_TC_DEBUG=$d CASES=$c timeout 200 python tools/fwd_time.py 2>&1 | grep -v Warning | grep -E "prof\]|us " | sort | uniq -c | sort -k2 | tail -30
  done
done > gpurun_out/prof_run.log 2>&1
cat gpurun_out/prof_run.log
