#!/bin/bash
out=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_r02g.log 2>&1
echo "pytest rc=$?" >> $out/pytest_r02g.log
grep -E "passed|failed|Error|assert" $out/pytest_r02g.log | tail -12
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
