#!/bin/bash
# Development helper: the library with the per-kernel %globaltimer stamps compiled in (-DCOEB_KERNEL_TRACE) ->
# coeb-slam_b200/build/variants/trace.so. Use: COEB_B200_LIB=.../trace.so COEB_KERNEL_TRACE=1 python tools/one_frame.py
set -e
cd "$(dirname "$0")/../coeb-slam_b200"
mkdir -p build/variants/trace
objs=""
for f in csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -DCOEB_KERNEL_TRACE "$@" -c $f -o build/variants/trace/$b.o &
  objs="$objs build/variants/trace/$b.o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/trace.so $objs -cudart static
echo build/variants/trace.so
