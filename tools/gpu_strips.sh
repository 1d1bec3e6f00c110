#!/bin/bash
# strips: the two-process IPC test, then the N-rank bench.   usage: gpurun --gpus N -- tools/gpu_strips.sh N tag
N=${1:-2}; tag=${2:-strips}
timeout 600 python -m pytest tests/test_strips_gpu.py -m gpu -q -x 2>&1 | tail -3
tools/gpu_multi.sh $N $tag | cut -c1-300 | tail -5
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n${N}_$tag.json").read().strip().splitlines()[-1])
print("value", d["value"], "one", d["one_frame_in_flight"]["value"])
print(json.dumps(d.get("strips"))[:1200])
PY
