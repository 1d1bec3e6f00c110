#!/bin/bash
# bench a set of tuning variants built by srp_b200.build.build_variant (run under gpurun)
for so in srp_b200/lib/libsrp_b200_*.so; do
  name=$(basename $so .so)
  SRP_B200_LIBRARY=$PWD/$so python bench.py --steps 20 --warmup 3 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(d['value'],1), {k:round(v,4) for k,v in d['stage_ms_per_frame'].items()})
"
done
