#!/bin/bash
# A/B builds of libbuddha.so: tools/build_ab.sh NAME [extra nvcc flags, e.g. -DBUDDHA_WARPS_PER_CTA=8]
# SRC=<dir with cudabrot_b200/csrc and include/> builds another source tree (e.g. an export of an older commit).
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=${SRC:-$ROOT}
mkdir -p $ROOT/tools/ab
cd $SRC/cudabrot_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -ccbin /usr/bin/g++ \
  --compiler-options "-fPIC -O2 -ffp-contract=off -fopenmp -Wall" "$@" -Xptxas -v -shared \
  -o $ROOT/tools/ab/$NAME.so buddha_api.cu -lgomp -ldl 2>&1 | grep -A2 "render_persistent_kernelILi0" | grep -E "registers|spill" || true
ls -la $ROOT/tools/ab/$NAME.so
