#!/bin/bash
# builds and runs the tcgen05 stand-alone test on the GPU box, always under a timeout (a wrong mbarrier
# protocol hangs instead of failing)
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr $TC_FLAGS -o /tmp/tc05_test tools/tc05_test.cu 2>&1 | grep -E "error" 
for rows in "$@"; do
  if [ "$rows" = "real" ]; then
    [ -f /tmp/nbr.bin ] || python tools/profile_conv.py --shapes 16x16 --reps 1 --dump /tmp/nbr.bin > /dev/null 2>&1
    timeout 120 /tmp/tc05_test /tmp/nbr.bin; echo "exit $?"
  else
    timeout 60 /tmp/tc05_test /nonexistent $rows; echo "exit $?"
  fi
done
