#!/bin/bash
# Round-end validation on a GPU box: full gpu test suite, smoke, bench (both arms), per-layer table, launch list + traffic.
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log; tail -3 gpurun_out/pytest_$tag.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-170 gpurun_out/bench_$tag.json
timeout 100 python tools/profile_layers.py 608 32 $tag > gpurun_out/layers_$tag.log 2>&1; head -1 gpurun_out/layers_$tag.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/metrics_$tag.csv python tools/ncu_target.py > gpurun_out/ncu_$tag.log 2>&1; tail -1 gpurun_out/ncu_$tag.log
