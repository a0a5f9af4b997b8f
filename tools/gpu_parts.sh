#!/bin/bash
out=gpurun_out
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --no-sub-configs --no-cpu-baseline > $out/p.json 2> $out/p.err
  python - <<PY
import json
d=json.load(open('gpurun_out/p.json'))
print('$1 step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['parity'])
PY
}
run default
DVDAGPU_NO_GRAPH=1 run nograph
DVDAGPU_PART_SECTORS=19000 run parts8
DVDAGPU_PART_SECTORS=12600 run parts12
