#!/bin/bash
# source-level ncu capture of ONE launch: tools/ncu_one.sh <tag> <cfg> <kernel regex> <skip>
tag=$1; cfg=$2; rx=$3; skip=${4:-0}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
  -f -o gpurun_out/ncu_$tag python tools/profile_target.py $cfg 1 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
