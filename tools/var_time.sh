#!/bin/bash
# tools/var_time.sh <variant names...>: forward timings of experimental library builds (exp/lib_<name>.so; "main" = the product build)
for v in "$@"; do
  if [ "$v" = main ]; then unset YC_LIB_PATH; else export YC_LIB_PATH=$PWD/exp/lib_$v.so; fi
  for d in ${DBG:-0}; do
    echo "== $v"
    YC_TC_DEBUG=$d CASES=${CASES:-s1,iaux,ibin} timeout 200 python tools/fwd_time.py 2>&1 | grep -v Warning
  done
done > gpurun_out/var_time.log 2>&1
cat gpurun_out/var_time.log
