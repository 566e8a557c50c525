#!/bin/bash
# Phase-knockout builds of libpcgc (tuning only, never shipped): PCGC_KNOCKOUT bit 0 = no halo fill, bit 1 = fragment LDS only
# for offset 0, bit 2 = weight LDS always from offset 0 (still issued), bit 3 = no MMAs (operands XORed into the result so
# that the loads stay).  Results are wrong by design; the time of what is left says which phase bounds
# conv_k3_octet_h2_kernel.  Only conv_h2.cu is recompiled; the rest links from the regular build.
#   here:        tools/knockout.sh 1 2 3 8 9 10 11
#   on the box:  PCGC_LIB=pcgcv2_b200/lib_ko/libpcgc_ko3.so python tools/profile_octet_h2.py --only octet_h2 --shapes 16x16
set -e
cd "$(dirname "$0")/../pcgcv2_b200/csrc"
mkdir -p ../lib_ko ../../build/pcgc_ko
for k in "$@"; do
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
      -DPCGC_KNOCKOUT=$k -c conv_h2.cu -o ../../build/pcgc_ko/conv_h2_$k.o 2>&1 | grep -E "rror" || true
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib_ko/libpcgc_ko$k.so \
      $(ls ../../build/pcgc/*.o | grep -v "/conv_h2.o") ../../build/pcgc_ko/conv_h2_$k.o -lcudart_static -lpthread -ldl -lrt ) &
done
wait
ls -la ../lib_ko/
