#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>: an experimental build of the library into exp/lib_<name>.so
# (git-ignored; picked up with YC_LIB_PATH=exp/lib_<name>.so).  Experiments only.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/exp/$name
mkdir -p $out
pids=()
for src in $root/yolo_continuous_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Wno-deprecated-gpu-targets \
    -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr "$@" -c $src -o $out/$(basename ${src%.cu}).o &
  pids+=($!)
done
for p in ${pids[@]}; do wait $p; done
/usr/local/cuda/bin/nvcc -shared -cudart static -o $root/exp/lib_$name.so $out/*.o
rm -rf $out
echo $root/exp/lib_$name.so
