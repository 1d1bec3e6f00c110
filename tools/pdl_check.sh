#!/bin/bash
# one GPU call: cfg3 bench with the programmatic-dependent-launch attribute off / on, then the
# parity suite with it on.  usage: gpurun -- tools/pdl_check.sh
mkdir -p gpurun_out
for p in 0 1; do
  SRP_B200_PDL=$p timeout 100 python bench.py --steps 30 --warmup 5 --cpu-seconds 0 > gpurun_out/bench_pdl$p.json 2> gpurun_out/bench_pdl$p.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_pdl$p.json").read().strip().splitlines()[-1])
    print("PDL=$p frames/s", round(d["value"],1), "ms", round(d["ms_per_step"],4), d["stage_ms_per_frame"], "e2e", round(d["e2e"]["value"],1), d["clocks"])
except Exception as e:
    print("PDL=$p bench failed", e)
PY
done
SRP_B200_PDL=1 timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_pdl1.log 2>&1
echo "pytest(PDL=1) exit $?"; tail -4 gpurun_out/pytest_pdl1.log
