#!/bin/bash
# bench.py under a list of environment settings: tools/envsweep.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 20 --warmup 3 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$var=$v', round(d['value'],1), {k:round(v,4) for k,v in d['stage_ms_per_frame'].items()})
"
done
