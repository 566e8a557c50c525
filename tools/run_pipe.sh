#!/bin/bash
# usage (on the GPU box): tools/run_pipe.sh   -- runs tools/bin/bench_pipe (built in the dev container with
# tools/build_pipe.sh "RUNP(...) RUNP(...)") on the finest decoder level's kernel map of the vox10 workload
cd "$(dirname "$0")/.."
[ -f /tmp/nbr.bin ] || python tools/profile_conv.py --shapes 16x16 --reps 1 --dump /tmp/nbr.bin > /dev/null
timeout 300 tools/bin/bench_pipe /tmp/nbr.bin
