#!/bin/bash
# parity suite, then the default bench line (headline + configs object)
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_r02b.log 2>&1
echo "pytest rc=$?" >> $out/pytest_r02b.log
tail -6 $out/pytest_r02b.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_r02b.json 2> $out/bench_r02b.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02b.json'))
print('headline', d['ms_per_step'], d['e2e']['ms_per_step'], 'parity', d['parity'], 'gen', d['generator_s'])
for k,v in d.get('configs',{}).items():
    print(k, round(v['ms_per_step'],3), round(v['e2e']['ms_per_step'],3), v.get('parity'), v['roofline_step']['frac'])
PY
grep -v "^$" $out/bench_r02b.err | tail -5
