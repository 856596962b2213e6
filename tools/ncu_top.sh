#!/bin/bash
# ncu --set full capture of the dominant kernel (conv_tc2_kernel<256, 4>) inside one warmed-up step.
# usage: tools/ncu_top.sh [out-name]   -> gpurun_out/<out-name>.ncu-rep
out=${1:-prof_top}
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k 'regex:conv_tc2_kernel<\(int\)256, \(int\)4>' -c 6 -f -o gpurun_out/$out python tools/ncu_target.py > gpurun_out/$out.log 2>&1
tail -3 gpurun_out/$out.log
