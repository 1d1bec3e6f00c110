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
# instruction counts / durations / issue utilisation of one cfg3 frame's kernels (a handful of metrics, few replays)
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --kernel-name regex:srpd --launch-skip 24 --launch-count 8 --csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_$tag.csv 2>/dev/null
python - <<PY
import csv
rows = [l for l in open("gpurun_out/ncu_$tag.csv") if l.startswith('"')]
last = None
for r in csv.DictReader(rows):
    k = r.get("Kernel Name", "?")[:40]
    if k != last:
        print(k); last = k
    print("   ", r.get("Metric Name"), r.get("Metric Value"))
PY
