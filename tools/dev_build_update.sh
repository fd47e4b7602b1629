#!/bin/bash
# Experiment builds of the update kernels: tools/dev_build_update.sh NAME "-DETB_SLICE_STAGES=8 ..." builds
# embeddingtables.jl_b200/lib/dev/libembtab_NAME.so (Float32 tables only: a quarter of the compile time); run with
# ETB_LIB_PATH=... .  The production library is always built by csrc/Makefile.
set -e
cd "$(dirname "$0")/../embeddingtables.jl_b200/csrc"
NAME=$1; shift
mkdir -p ../build/dev ../lib/dev
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr -cudart static -DETB_DEV_F32_ONLY "$@" -c etb_update.cu -o ../build/dev/etb_update_$NAME.o 2>/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../lib/dev/libembtab_$NAME.so \
  ../build/etb_runtime.o ../build/etb_lookup.o ../build/etb_index.o ../build/dev/etb_update_$NAME.o -ldl
echo built $NAME
