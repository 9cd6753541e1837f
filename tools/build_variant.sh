#!/bin/bash
# Development helper: builds a variant of the library with extra -D flags for one source file.
#   tools/build_variant.sh NAME file.cu -DFOO=1 ...   ->  coeb-slam_b200/build/variants/NAME.so   (use with COEB_B200_LIB=...)
set -e
cd "$(dirname "$0")/../coeb-slam_b200"
name=$1; src=$2; shift 2
make -s >/dev/null
mkdir -p build/variants
base=$(basename "$src" .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v "$@" -c csrc/$src -o build/variants/$name.$base.o 2> build/variants/$name.ptxas.log
objs=""
for o in build/*.o; do if [ "$(basename $o .o)" = "$base" ]; then objs="$objs build/variants/$name.$base.o"; else objs="$objs $o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $objs -cudart static
grep -A2 "${KERNEL:-fast_kernel}" build/variants/$name.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
