#!/bin/bash
# usage (dev container): tools/build_pipe.sh "RUNP(16,16,false,4,3,2,7) RUNP(16,4,true,2,3,2,7) ..." [output name]
# RUNP(CIN, COUT, NT, RG, D, MINB, OPT): one conv_k3_pipe_kernel variant per entry
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
echo "$1" > /tmp/variants_pipe.h
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -o tools/bin/${2:-bench_pipe} tools/bench_pipe.cu
