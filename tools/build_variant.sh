#!/bin/bash
# Build a library variant for A/B runs: tools/build_variant.sh name [-DFLAG=..]...  -> variants/name.so (loaded via DTA_B200_LIB)
# SRC=dir overrides the source directory (e.g. a checkout of an older commit's csrc).
set -e
NAME=$1; shift
SRC=${SRC:-deeptreeattention_b200/csrc}
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" $SRC/dta_api.cu $SRC/dta_blocks_api.cu -o variants/$NAME.so
echo built variants/$NAME.so
