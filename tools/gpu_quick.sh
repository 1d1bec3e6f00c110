#!/bin/bash
# quick GPU iteration: parity suite, short bench, source-level ncu capture of the tile kernel.
# usage: gpurun -- tools/gpu_quick.sh <tag> [kernel regex]
tag=${1:-q}; rx=${2:-srpdTileKernel}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_$tag.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("frames/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), d["stage_ms_per_frame"], "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
except Exception as e:
    print("bench failed", e)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 \
  -f -o gpurun_out/ncu_$tag python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
