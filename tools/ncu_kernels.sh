#!/bin/bash
# ncu --set full captures (2 launches each) of the kernels DESIGN.md discusses, inside one warmed-up step (tools/ncu_target.py),
# plus the launch list of the whole step with DRAM bytes / tensor-pipe % per launch.   usage: tools/ncu_kernels.sh <tag>
tag=${1:-r02}
cap() {  # name regex
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -c 2 -f -o gpurun_out/${tag}_$1 python tools/ncu_target.py > gpurun_out/${tag}_$1.log 2>&1
  tail -1 gpurun_out/${tag}_$1.log
}
cap tc2_256 'conv_tc2_kernel<\(int\)256, \(int\)8>'
cap tc2_128 'conv_tc2_kernel<\(int\)128, \(int\)8>'
cap lean64 'conv_tc_kernel<\(int\)64, \(int\)64, \(bool\)0, \(int\)4, \(bool\)1>'
cap stem 'stem_tc_kernel'
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --profile-from-start off --kernel-name-base demangled --csv --log-file gpurun_out/${tag}_launches_metrics.csv \
    python tools/ncu_target.py > gpurun_out/${tag}_launches.log 2>&1
tail -1 gpurun_out/${tag}_launches.log
