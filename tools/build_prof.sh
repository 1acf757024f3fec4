#!/bin/sh
# instrumented build of the library (per-phase clock counts printed by CTA 0): tools/libslod_prof.so
cd "$(dirname "$0")/../dealii-slod_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DSLOD_PHASE_CLOCKS -o ../../tools/libslod_prof.so kernels.cu capi.cu
