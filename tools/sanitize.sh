#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the kernels of the hot path: one small forward + decode/NMS per
# precision mode and per forced tcgen05 plan family (single CTA / CTA pair, slab / per-thread epilogue, lean + resident W).
# usage (GPU box): tools/sanitize.sh [tag]   ->  gpurun_out/sanitize_<tag>.md + per-run logs
tag=${1:-r02}
out=gpurun_out/sanitize_$tag.md
mkdir -p gpurun_out
echo "# compute-sanitizer summary ($tag): tools/sanitize.sh, 64x64 batch 2 forward + decode/NMS incl. overflow path" > $out
echo '| tool | precision | forced plan (Y4_FORCE) | target | sanitizer summary |' >> $out
echo '|---|---|---|---|---|' >> $out
run() {   # tool prec force
  local log=gpurun_out/sanitize_${tag}_$1_$2_$(echo "$3" | tr ',' '_').log
  if [ -n "$3" ]; then export Y4_FORCE="$3"; else unset Y4_FORCE; fi
  Y4_AUTOTUNE=${4:-0} timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool $1 --print-limit 10 python tools/sanitize_target.py $2 > $log 2>&1
  local ok=$(grep -c 'SANITIZE TARGET OK' $log)
  local summ=$(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1 | sed 's/=========//')
  echo "| $1 | $2 | ${3:-default plans} | $([ $ok -ge 1 ] && echo ran || echo FAILED) | ${summ:-none printed} |" >> $out
}
if [ "$2" = "decode" ]; then      # tools/sanitize.sh <tag> decode: only the decode / NMS kernels (fast)
  for tool in memcheck racecheck synccheck; do run $tool decode ""; done
  cat $out
  exit 0
fi
for tool in memcheck racecheck synccheck; do
  run $tool fp16 ""
  run $tool fp16 "128,224,0,1,1,4,0,32,0,0,0" 1
  run $tool fp16 "256,224,0,2,1,8,0,32,1,0,0" 1
  run $tool fp16 "64,75,0,1,1,4,1,64,0,1,0" 1
  run $tool fp16x3 ""
  run $tool fp32 ""
done
cat $out
