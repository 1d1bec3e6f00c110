#!/bin/bash
# parity suite + short bench + all configs + cfg4 launch list
tag=${1:-chk}
tools/gpu_check.sh $tag
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 600 ncu --metrics $M --clock-control none --kernel-name regex:srpd --csv python tools/profile_target.py cfg4 1 > gpurun_out/l4_$tag.csv 2> gpurun_out/l4_$tag.err
python - <<PY
import csv, collections
rows = [l for l in open("gpurun_out/l4_$tag.csv") if l.startswith('"')]
agg = collections.OrderedDict()
for r in csv.DictReader(rows):
    agg.setdefault((r["ID"], r["Kernel Name"][:44]), {})[r["Metric Name"]] = r["Metric Value"]
for (i, k), m in agg.items():
    f = lambda n: float(m.get(n, "0").replace(",", ""))
    if f('gpu__time_duration.sum') > 30e3:
        print(f"{k:46s} {f('gpu__time_duration.sum')/1e3:9.1f} us inst {f('smsp__inst_executed.sum')/1e6:8.2f} M occ {f('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} thr {f('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f}")
PY
rm -f gpurun_out/l4_$tag.csv
