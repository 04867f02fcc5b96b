#!/bin/bash
# Builds launch-shape variants of libso100_b200.so into build/ for A/B timing on the GPU box (tools/bench_variants.sh).
set -e
cd "$(dirname "$0")/.."
mkdir -p build
build() { # tag block minblocks sync [extra flags]
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
    -DSO100_BLOCK=$2 -DSO100_MINBLOCKS=$3 -DSO100_SYNC=$4 $5 -Xptxas -v \
    -o build/libso100_$1.so so100_mujoco_rl_b200/csrc/so100_b200.cu 2>&1 | grep -A2 "step_kernelILi1E" | grep -E "registers|spill" | tr '\n' ' '
  echo " <- $1"
}
build b256s 256 2 1 &
build skew2k 256 2 1 "-DSO100_SKEW=2000" &
build skew4k 256 2 1 "-DSO100_SKEW=4000" &
build skew7k 256 2 1 "-DSO100_SKEW=7000" &
wait
build b128s4 128 4 1 &
build b128s4skew 128 4 1 "-DSO100_SKEW=3500" &
wait
