#!/bin/bash
# Build variants/lib_<name>.so = the in-tree objects with ONE kernel-instantiation unit recompiled under extra -D flags
# (same-box A/B of kernel tuning macros without rebuilding everything).  usage: tools/build_variant.sh name unit "flags"
#   e.g. tools/build_variant.sh tpt1_1024 inst_mackey_glass10 "-DSDDE_TUNE_TPT_SCALAR=1 -DSDDE_BLOCK_SCALAR=1024"
set -e
NAME=$1; UNIT=$2; FLAGS=$3
cd "$(dirname "$0")/../stochasticdelaydiffeq.jl_b200/csrc"
mkdir -p ../../variants/obj
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  $FLAGS -c $UNIT.cu -o ../../variants/obj/${UNIT}_$NAME.o
OBJS=$(ls build/*.o | grep -v "build/$UNIT.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$NAME.so $OBJS ../../variants/obj/${UNIT}_$NAME.o \
  -lnvrtc -ldl -Xlinker --exclude-libs,ALL
cuobjdump -res-usage ../../variants/obj/${UNIT}_$NAME.o 2>&1 | grep -A1 "ELi0ELi0ELb0E" | grep -o "REG:[0-9]* STACK:[0-9]*" | head -1
