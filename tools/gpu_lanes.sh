#!/bin/bash
# GPU iteration on frames in flight: parity suite, then the bench with 1..4 lanes.
# usage: gpurun -- tools/gpu_lanes.sh <tag>
tag=${1:-lanes}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_$tag.log | cut -c1-300
for lanes in 1 2 3 4; do
  timeout 300 python bench.py --steps 40 --warmup 3 --cpu-seconds 0 --lanes $lanes > gpurun_out/bench_${tag}_l$lanes.json 2> gpurun_out/bench_${tag}_l$lanes.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_l$lanes.json").read().strip().splitlines()[-1])
    print("lanes", $lanes, "frames/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "one at a time", round(d["one_frame_in_flight"]["value"], 1),
          d["stage_ms_per_frame"], "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], "equal", d.get("frames_in_flight_planes_equal"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${tag}_l$lanes.err").read()[-1500:])
PY
done
