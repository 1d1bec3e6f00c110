#!/bin/bash
# GPU iteration: parity suite, short bench, all-config timings.   usage: gpurun -- tools/gpu_check.sh <tag>
tag=${1:-chk}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_$tag.log | cut -c1-400
timeout 300 python bench.py --steps 40 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("frames/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "one at a time", round(d["one_frame_in_flight"]["value"], 1),
          d["stage_ms_per_frame"], "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], "equal", d.get("frames_in_flight_planes_equal"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$tag.err").read()[-1500:])
PY
timeout 400 python tools/bench_all.py > gpurun_out/bench_all_$tag.log 2>&1
cp gpurun_out/bench_all.json gpurun_out/bench_all_$tag.json 2>/dev/null
python - <<PY
import json
try:
    for r in json.load(open("gpurun_out/bench_all_$tag.json"))["results"]:
        print(r["scene"][:40], {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("scene",)})
except Exception as e:
    print("bench_all failed", e)
PY
