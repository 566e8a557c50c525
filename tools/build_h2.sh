#!/bin/bash
# usage (dev container): tools/build_h2.sh "RUNH(16,16,false,4,2,8,2) RUNH(16,4,true,2,2,8,2) ..." [output name]
# RUNH(CIN, COUT, NT, RG, D, WARPS, MINB): one conv_k3_h2_kernel variant per entry
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
echo "$1" > /tmp/variants_h2.h
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xptxas -v -o tools/bin/${2:-bench_h2} tools/bench_h2.cu 2>&1 | grep -A1 "conv_k3_h2" | grep -v "^--" | paste - - | sed 's/ptxas info    : //g' | awk '{print}' | cut -c1-400
