#!/bin/bash
# one GPU call: parity suite, then the cfg3 bench under the environment settings given as
# arguments ("-" = defaults).  usage: gpurun -- tools/iter_check.sh - SRP_B200_CKPT_ASIDE=0 ...
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_iter.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_iter.log
i=0
for setting in "$@"; do
  i=$((i+1))
  envs=""; [ "$setting" != "-" ] && envs="$setting"
  env $envs timeout 100 python bench.py --steps 30 --warmup 5 --cpu-seconds 0 > gpurun_out/bench_iter$i.json 2> gpurun_out/bench_iter$i.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_iter$i.json").read().strip().splitlines()[-1])
    print("[$setting] frames/s", round(d["value"],1), "ms", round(d["ms_per_step"],4), d["stage_ms_per_frame"], "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], d["clocks"])
except Exception as e:
    print("[$setting] bench failed", e)
PY
done
# the heavy-clipping variant of cfg3 (large triangles near the camera) with the checkpoint pre-pass
# beside the binning kernels / in the main stream
for c in 1 0; do
  echo "cfg3_r12 CKPT_ASIDE=$c"; SRP_B200_CKPT_ASIDE=$c timeout 100 python tools/bench_all.py cfg3_r12 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(round(d['ms_per_frame'],4), d['stage_ms_per_draw'])
"
done
