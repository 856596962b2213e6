#!/bin/bash
# Round-end validation on a GPU box: full gpu test suite, smoke, bench (both arms), per-layer table, launch list + traffic.
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log; tail -3 gpurun_out/pytest_$tag.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-170 gpurun_out/bench_$tag.json
timeout 100 python tools/profile_layers.py 608 32 $tag > gpurun_out/layers_$tag.log 2>&1; head -1 gpurun_out/layers_$tag.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/metrics_$tag.csv python tools/ncu_target.py > gpurun_out/ncu_$tag.log 2>&1; tail -1 gpurun_out/ncu_$tag.log
# config 4 (decode + NMS isolation), its --set full capture, the dominant conv kernels' --set full captures, decode sanitizer
timeout 200 python tools/bench_decode_nms.py 608 32 > gpurun_out/decode_nms_$tag.log 2>&1; tail -1 gpurun_out/decode_nms_$tag.log | cut -c1-120
cp gpurun_out/decode_nms.json gpurun_out/decode_nms_$tag.json    # the ncu capture below re-runs the tool and overwrites decode_nms.json with profiler-perturbed times
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"nms_image|decode_filter" -s 6 -c 2 -f -o gpurun_out/${tag}_decode python tools/bench_decode_nms.py 608 32 > /dev/null 2>&1
for k in 'conv_tc2_kernel<\(int\)256, \(int\)8>' 'conv_tc2_kernel<\(int\)256, \(int\)16>'; do
  n=$(echo "$k" | tr -dc '0-9' | tail -c 5)
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k "regex:$k" -c 2 -f -o gpurun_out/${tag}_tc2_$n python tools/ncu_target.py > /dev/null 2>&1
done
ls gpurun_out/${tag}_*.ncu-rep
timeout 600 bash tools/sanitize.sh ${tag} decode | tail -4
timeout 300 python tools/bench_sweep.py 16 > gpurun_out/sweep_$tag.log 2>&1; cut -c1-140 gpurun_out/sweep_$tag.log | tail -5
