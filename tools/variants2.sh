#!/bin/bash
# parity of a few scenes + bench for each variant library (run under gpurun)
for so in srp_b200/lib/libsrp_b200*.so; do
  name=$(basename $so .so)
  SRP_B200_LIBRARY=$PWD/$so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "synthetic_scene or full_size_configs" 2>&1 | tail -1
  for i in 1 2; do
  SRP_B200_LIBRARY=$PWD/$so python bench.py --steps 40 --warmup 3 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(d['value'],1), round(d['one_frame_in_flight']['value'],1), {k:round(v,4) for k,v in d['stage_ms_per_frame'].items()})
"
  done
done
