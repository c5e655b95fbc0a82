#!/bin/bash
# Developer tool: build a variant of the float instantiation with extra -D flags into tools/bin/lib_<name>.so
# usage: tools/build_variant.sh <name> [-DFLAG=VAL ...]
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
CS=$HERE/papteam_opticalflow_b200/csrc
OD=$HERE/tools/bin/obj_$NAME
mkdir -p $OD
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-Wall,-Wno-unused-function -diag-suppress 128 "$@" -c $CS/inst_f32.cu -o $OD/inst_f32.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared \
    -o $HERE/tools/bin/lib_$NAME.so $CS/build/inst_f64.o $OD/inst_f32.o $CS/build/cabi.o
echo built tools/bin/lib_$NAME.so
