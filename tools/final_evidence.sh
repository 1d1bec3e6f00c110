#!/bin/bash
# round-end evidence in one GPU call: the driver's bench line (with the CPU baseline), the
# reference arm, the ncu launch list, one `--set full` capture of a frame's kernels, all configs.
# usage: gpurun -- tools/final_evidence.sh
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.json; echo
timeout 120 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 300 gpurun_out/bench_ref.json; echo
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:srpd -s 27 -c 9 -f -o gpurun_out/full \
  python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/full_ncu.log 2>&1
tail -2 gpurun_out/full_ncu.log
timeout 200 python tools/bench_all.py > gpurun_out/bench_all.log 2>&1
tail -3 gpurun_out/bench_all.log | cut -c1-300
ls -la gpurun_out
