#!/bin/bash
# Builds launch-shape variants of libso100_b200.so into build/ for A/B timing on the GPU box (tools/bench_variants.sh).
set -e
cd "$(dirname "$0")/.."
mkdir -p build
rm -f build/libso100_*.so
build() { # tag block minblocks sync [extra flags]
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
    -DSO100_BLOCK=$2 -DSO100_MINBLOCKS=$3 -DSO100_SYNC=$4 $5 -Xptxas -v \
    -o build/libso100_$1.so so100_mujoco_rl_b200/csrc/so100_b200.cu 2>&1 | grep -A2 "step_kernelILi1ELb1" | grep -E "registers|spill" | tr '\n' ' '
  echo " <- $1"
}
build b256s 256 2 1 &
build b256 256 2 0 &
build b128x4s 128 4 1 &
build b128x4 128 4 0 &
wait
build b512s 512 1 1 &
build b64x8 64 8 0 &
wait
