#!/bin/bash
# usage: tools/exp_force.sh tag "bn,budgetKB,patch,group,epi,nepi,bres,gw[,pair]" [size batch]  -> gpurun_out/layers_<tag>.json (per-layer times with that plan forced)
tag=$1; force=$2; size=${3:-608}; batch=${4:-32}
Y4_FORCE="$force" timeout 200 python tools/profile_layers.py $size $batch $tag > gpurun_out/layers_$tag.log 2>&1
head -1 gpurun_out/layers_$tag.log
