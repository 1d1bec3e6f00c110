#!/bin/bash
tag=${1:-lanes}
for lanes in 4 6 8; do
 for steps in 20 40; do
  timeout 300 python bench.py --steps $steps --warmup 5 --cpu-seconds 0 --lanes $lanes > gpurun_out/bench_${tag}_l$lanes.json 2> gpurun_out/bench_${tag}_l$lanes.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_l$lanes.json").read().strip().splitlines()[-1])
    print("lanes", $lanes, "steps", $steps, "frames/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "one at a time", round(d["one_frame_in_flight"]["value"], 1), "equal", d.get("frames_in_flight_planes_equal"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${tag}_l$lanes.err").read()[-1500:])
PY
 done
done
