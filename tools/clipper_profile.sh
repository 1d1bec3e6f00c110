#!/bin/bash
# source-level ncu capture of one clipper-pass launch (srpdGeomKernel<*, true>) of a cfg3 frame.
# usage: gpurun -- tools/clipper_profile.sh
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:srpdGeomKernel -s 7 -c 1 \
  -f -o gpurun_out/clipper python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/clipper_ncu.log 2>&1
tail -3 gpurun_out/clipper_ncu.log
ls -la gpurun_out/
