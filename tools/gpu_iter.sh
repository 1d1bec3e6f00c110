#!/bin/bash
# development loop on the GPU box: parity tests, a short bench, and the tile/geometry kernels'
# instruction counts + durations (ncu, two metrics only).  usage: gpurun -- tools/gpu_iter.sh [tag]
tag=${1:-iter}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print("frames/s", round(d["value"],1), "ms", round(d["ms_per_step"],4), d["stage_ms_per_frame"], "e2e", round(d["e2e"]["value"],1), d["e2e"]["variants"], d["clocks"])
PY
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name regex:srpd --launch-skip 27 --launch-count 9 --csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_$tag.csv 2>/dev/null
python - <<PY
import csv
rows=[l for l in open("gpurun_out/ncu_$tag.csv") if l.startswith('"')]
for r in csv.DictReader(rows):
    print(r.get('Kernel Name','?')[:34], r.get('Metric Name'), r.get('Metric Value'))
PY
