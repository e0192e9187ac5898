#!/bin/bash
# Timing experiments on reg_forward_tc: rebuild reg_tc_kernels.cu with one piece of the sweep removed (results are garbage,
# only the kernel time matters) and link a side library deepfbsdejsolvers_b200/libfbsdej_ablate_K.so; run with
#   FBSDEJ_LIB=deepfbsdejsolvers_b200/libfbsdej_ablate_K.so python bench.py --no-cpu-baseline --no-e2e
set -e
cd "$(dirname "$0")/../deepfbsdejsolvers_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
for k in "$@"; do  # 1-5: forward, 11-14: adjoint (11 no weight-gradient GEMMs, 12 no layer GEMMs, 13 no tanh, 14 no lo-tile stores)
  ( nvcc $FLAGS -DFBSDEJ_ABLATE=$k -c reg_tc_kernels.cu -o build/reg_tc_ablate_$k.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libfbsdej_ablate_$k.so build/api.o build/sim_kernels.o build/util_kernels.o \
      build/pricing_kernels.o build/reg_tc_ablate_$k.o build/mfg_kernels.o build/mfg_tc_kernels.o -lcudart ) &
done
wait
