#!/bin/bash
tag=${1:-c4t}
tools/ncu_one.sh ${tag}_linetile cfg4 srpdTileKernel 1
tools/ncu_one.sh ${tag}_tritile cfg4 srpdTileKernel 0
tools/ncu_one.sh ${tag}_trigeom cfg4 srpdGeomKernel 0
ls -la gpurun_out/ncu_${tag}_*
