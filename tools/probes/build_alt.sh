#!/bin/bash
# build_alt.sh NAME "-DFLAG=..." : lib/libfa_alt_NAME.so = the product objects with csrc/fa_tc_fwd.cu recompiled under extra flags
# (A/B experiments inside one gpurun call: FA_SM100A_LIB=.../libfa_alt_NAME.so python tools/probes/ab_fwd.py)
set -e
cd "$(dirname "$0")/../../flashattention.jl_b200"
name=$1; shift
src=${SRC:-fa_tc_fwd}
mkdir -p build/alt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" -c csrc/$src.cu -o build/alt/${src}_$name.o 2> build/alt/${src}_$name.log
objs=$(ls build/*.o | grep -v "build/$src.o" | grep -v probe)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/libfa_alt_$name.so $objs build/alt/${src}_$name.o -cudart static -ldl
grep -A2 "${GREP:-tc_fwd_kernelILi128ELi1ELi2ELi1}" build/alt/${src}_$name.log | tail -2
