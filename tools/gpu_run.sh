#!/bin/bash
# one GPU call: parity suite, driver-style bench line, all-config timings, host-link ceiling, topology.
# usage: gpurun -- tools/gpu_run.sh <tag> [pytest-args...]
tag=${1:-run}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short "$@" > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_$tag.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_$tag.json; echo; tail -5 gpurun_out/bench_$tag.err
timeout 400 python tools/bench_all.py > gpurun_out/bench_all_$tag.log 2>&1
cp gpurun_out/bench_all.json gpurun_out/bench_all_$tag.json 2>/dev/null
python - <<PY
import json
try:
    for r in json.load(open("gpurun_out/bench_all_$tag.json"))["results"]:
        print(r["scene"][:40], {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != "scene"})
except Exception as e:
    print("bench_all failed", e)
PY
timeout 120 python tools/pcie_ceiling.py > gpurun_out/pcie_$tag.log 2>&1; head -1 gpurun_out/pcie_$tag.log
(nvidia-smi topo -m; numactl -H; lscpu | head -30; nproc) > gpurun_out/topology_$tag.txt 2>&1
