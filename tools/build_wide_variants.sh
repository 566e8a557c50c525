#!/bin/bash
# Tuning variants of the tcgen05 kernel as separate libraries (build/variants/libpcgc_<tag>.so, selected with PCGC_LIB=...):
#   tools/build_wide_variants.sh "<tag> <nvcc -D flags>" ...
set -e
cd "$(dirname "$0")/../pcgcv2_b200/csrc"
mkdir -p ../../tools/bin/variants
for spec in "$@"; do
  tag=${spec%% *}; flags=${spec#* }
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c conv_wide.cu -o ../../tools/bin/variants/conv_wide_$tag.o
  objs=$(ls ../../build/pcgc/*.o | grep -v conv_wide.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/variants/libpcgc_$tag.so $objs ../../tools/bin/variants/conv_wide_$tag.o -lcudart_static -lpthread -ldl -lrt
  echo built $tag
done
