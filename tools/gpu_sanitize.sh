#!/bin/bash
# compute-sanitizer over the synthetic parity scenes (memcheck, then racecheck on a few)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "synthetic_scene or frames_in_flight or large_line_draw" > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/memcheck_pytest.log | cut -c1-200; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck.log; tail -5 gpurun_out/memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
  python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "synthetic_scene or binned_and_direct or large_line_draw" > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck exit $?"; tail -3 gpurun_out/racecheck_pytest.log | cut -c1-200; grep -c "hazard" gpurun_out/racecheck.log; tail -5 gpurun_out/racecheck.log | cut -c1-200
timeout 300 python bench.py --steps 40 --warmup 3 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('in flight', round(d['value'],1), 'one', round(d['one_frame_in_flight']['value'],1), {k:round(v,4) for k,v in d['stage_ms_per_frame'].items()})
"
