#!/bin/bash
# cfg4: launch list + source-level captures of the 1M-line draw's geometry and tile kernels
tag=${1:-c4}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
c=cfg4
timeout 600 ncu --metrics $M --clock-control none --kernel-name regex:srpd --csv python tools/profile_target.py $c 2 > gpurun_out/launches_${c}_$tag.csv 2> gpurun_out/launches_${c}_$tag.err
python - <<PY
import csv, collections
rows = [l for l in open("gpurun_out/launches_${c}_$tag.csv") if l.startswith('"')]
agg = collections.OrderedDict()
for r in csv.DictReader(rows):
    key = (r["ID"], r["Kernel Name"][:44])
    agg.setdefault(key, {})[r["Metric Name"]] = r["Metric Value"]
print("== $c")
items = list(agg.items())
half = len(items) // 2
for (i, k), m in items[half:]:
    f = lambda n: float(m.get(n, "0").replace(",", ""))
    print(f"{k:46s} {f('gpu__time_duration.sum')/1e3:9.1f} us  rd {f('dram__bytes_read.sum')/1e6:8.1f} MB wr {f('dram__bytes_write.sum')/1e6:8.1f} MB  inst {f('smsp__inst_executed.sum')/1e6:8.2f} M  occ {f('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}  issue {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} thr {f('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f}")
PY
rm -f gpurun_out/launches_${c}_$tag.csv
tools/ncu_one.sh ${tag}_linegeom cfg4 'srpdGeomKernel' 2
tools/ncu_one.sh ${tag}_linetile cfg4 'srpdTileKernelILi1' 0
tools/ncu_one.sh ${tag}_tri cfg4 'srpdGeomKernel' 0
ls -la gpurun_out/ncu_${tag}_*
