#!/bin/bash
# ncu --set full capture of the merged first-stage kernel (conv_k3_octet_h2_kernel<16, 8>) in its production context
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_k3_octet_h2_kernel<\(int\)16, \(int\)8' -c 1 -f -o gpurun_out/r02_prof_merged python tools/one_frame.py
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_k3_ones_from_parent' -c 1 -f -o gpurun_out/r02_prof_ones python tools/one_frame.py
