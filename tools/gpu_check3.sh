#!/bin/bash
tag=${1:-chk}
tools/gpu_check2.sh $tag
tools/ncu_one.sh ${tag}_linetile cfg4 'srpdTileKernel<\(int\)1' 0
tools/ncu_one.sh ${tag}_tritile cfg4 'srpdTileKernel<\(int\)0' 0
ls -la gpurun_out/ncu_${tag}_*
