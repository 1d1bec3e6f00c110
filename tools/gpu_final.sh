#!/bin/bash
# last check of a round: smoke(), the whole GPU suite, the driver's two bench commands
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_final.log | cut -c1-300
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/final_ref.json 2>/dev/null; tail -c 300 gpurun_out/final_ref.json; echo
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "one", round(d["one_frame_in_flight"]["value"], 1), d["stage_ms_per_frame"], "e2e", round(d["e2e"]["value"], 1),
      "launches", d["gpu_launches"], "parity", d["parity"]["planes_equal"], "clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"])
PY
