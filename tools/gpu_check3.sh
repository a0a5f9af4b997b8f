#!/bin/bash
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_r02d.log 2>&1
echo "pytest rc=$?" >> $out/pytest_r02d.log
tail -25 $out/pytest_r02d.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-sub-configs > $out/bench_r02d.json 2> $out/bench_r02d.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_r02d.json'))
    print('headline', d['ms_per_step'], d['e2e']['ms_per_step'], 'parity', d['parity'], d['gpu_launches'])
    print(d['kernel_ms_per_step']); print(d['stage_ms_per_step'])
except Exception as e: print('ERR', e)
PY
tail -5 $out/bench_r02d.err
timeout 120 python tools/trace_decode.py 600 > $out/trace_r02d.txt 2>&1; tail -50 $out/trace_r02d.txt
