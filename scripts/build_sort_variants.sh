#!/bin/bash
# Builds A/B variants of the onesweep sort (scripts/sort_bench.cu) into gpurun_variants/.
# usage: scripts/build_sort_variants.sh name "-DFLAG=.. -DFLAG=.." [name flags ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
while [ $# -ge 2 ]; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a $2 \
       scripts/sort_bench.cu cuspatial_b200/csrc/radix_sort.cu -o gpurun_variants/$1 &
  shift 2
done
wait
