#!/bin/bash
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_r02i.log 2>&1
echo "pytest rc=$?" >> $out/pytest_r02i.log
grep -E "passed|failed|FAILED|Error|assert " $out/pytest_r02i.log | tail -8
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --no-sub-configs --no-cpu-baseline > $out/p.json 2> $out/p.err
  python - <<PY
import json
d=json.load(open('gpurun_out/p.json'))
print('$1 step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['parity'])
print({k:v for k,v in d['kernel_ms_per_step'].items()})
PY
}
run pdl
DVDAGPU_NO_PDL=1 run nopdl
timeout 120 python tools/trace_decode.py 600 > $out/trace_r02i.txt 2>&1; tail -45 $out/trace_r02i.txt | cut -c1-110
