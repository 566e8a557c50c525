#!/bin/bash
# usage: tools/run_variants.sh "RUN(16,16,4,3,1) RUN(16,16,2,3,1) ..."   (runs on the GPU box)
set -e
cd "$(dirname "$0")/.."
[ -f /tmp/nbr.bin ] || python tools/profile_conv.py --shapes 16x16 --reps 1 --dump /tmp/nbr.bin > /dev/null
echo "$1" > /tmp/variants.h
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -o /tmp/bench_variants tools/bench_conv_variants.cu
/tmp/bench_variants /tmp/nbr.bin
