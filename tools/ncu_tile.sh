#!/bin/bash
# source-level ncu capture of one tile-kernel launch of a cfg3 frame.  usage: gpurun -- tools/ncu_tile.sh [tag] [kernel regex]
tag=${1:-tile}; rx=${2:-srpdTileKernel}
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 \
  -f -o gpurun_out/ncu_$tag python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
ls -la gpurun_out/ncu_$tag.ncu-rep
