#!/bin/bash
# usage (on the GPU box): tools/run_h2.sh [binary name]
cd "$(dirname "$0")/.."
[ -f /tmp/nbr.bin ] || python tools/profile_conv.py --shapes 16x16 --reps 1 --dump /tmp/nbr.bin > /dev/null
timeout 300 tools/bin/${1:-bench_h2} /tmp/nbr.bin $2
