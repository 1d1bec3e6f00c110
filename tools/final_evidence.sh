#!/bin/bash
# round evidence in one GPU call (1 GPU): the driver's bench line (with parity + CPU baseline), the
# reference arm, the ncu launch list and one `--set full` capture of a cfg3 frame's kernels, full
# captures of the cfg5-batch kernels and of cfg4's binning + tile kernels, launch lists of cfg4 / cfg5 /
# cfg2, all-config timings.   usage: gpurun -- tools/final_evidence.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg3_n1.json 2> gpurun_out/${tag}_bench_cfg3_n1.err
tail -c 400 gpurun_out/${tag}_bench_cfg3_n1.json; echo
timeout 200 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${tag}_bench_cfg3_reference_arm.json 2>/dev/null
tail -c 300 gpurun_out/${tag}_bench_cfg3_reference_arm.json; echo
# the CPU figures that stand next to the other configs (same arm, other workloads)
: > gpurun_out/${tag}_reference_arm_all_configs.jsonl
for w in cfg1 cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --impl reference --workload $w --steps 3 --warmup 1 2>/dev/null | tail -1 >> gpurun_out/${tag}_reference_arm_all_configs.jsonl
done
cut -c1-160 gpurun_out/${tag}_reference_arm_all_configs.jsonl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_cfg3_launches.csv \
  python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:srpd -s 24 -c 8 -f -o gpurun_out/${tag}_cfg3_full \
  python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/${tag}_cfg3_full.log 2>&1
tail -1 gpurun_out/${tag}_cfg3_full.log
python tools/ncu_summary.py gpurun_out/${tag}_cfg3_full.ncu-rep gpurun_out/${tag}_cfg3_launches.csv gpurun_out/${tag}_cfg3_ncu_summary.md "Round 2 -- cfg3 (1M-triangle shell, 3840x2160): ncu evidence"
timeout 300 ncu --set full --clock-control none -k regex:srpd -c 5 -f -o gpurun_out/${tag}_cfg5_batch_full \
  python tools/profile_target.py cfg5 1 > gpurun_out/${tag}_cfg5_batch_full.log 2>&1
tail -1 gpurun_out/${tag}_cfg5_batch_full.log
python tools/ncu_summary.py gpurun_out/${tag}_cfg5_batch_full.ncu-rep /dev/null gpurun_out/${tag}_cfg5_batch_ncu_summary.md "Round 2 -- cfg5 batch (1024 teapot frames 1024x1024, one srpB200DrawBatch): ncu --set full"
rm -f gpurun_out/${tag}_cfg5_batch_full.ncu-rep      # (gpurun_out/ travels back only below 64 MiB: the summaries are what is kept)
timeout 400 ncu --set full --clock-control none -k 'regex:srpdBin|srpdTile|srpdBatchOrder' -c 6 -f -o gpurun_out/${tag}_cfg4_bin_full \
  python tools/profile_target.py cfg4 1 > gpurun_out/${tag}_cfg4_bin_full.log 2>&1
tail -1 gpurun_out/${tag}_cfg4_bin_full.log
python tools/ncu_summary.py gpurun_out/${tag}_cfg4_bin_full.ncu-rep /dev/null gpurun_out/${tag}_cfg4_bin_ncu_summary.md "Round 2 -- cfg4 first draw (10 M sub-pixel triangles, stencil + scissor): order, binning and tile kernels, ncu --set full"
rm -f gpurun_out/${tag}_cfg4_bin_full.ncu-rep
tools/ncu_configs.sh $tag > gpurun_out/${tag}_launch_lists.txt 2>&1
rm -f gpurun_out/launches_*_${tag}.csv
timeout 400 python tools/bench_all.py > gpurun_out/${tag}_bench_all.log 2>&1
cp gpurun_out/bench_all.json gpurun_out/${tag}_bench_all_configs.json 2>/dev/null
ls -la gpurun_out | grep ${tag}_
