#!/bin/bash
out=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_r02k.json 2> $out/bench_r02k.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02k.json'))
print('headline', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['parity'], 'gen', round(d['generator_s'],1), 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'step frac', round(d['roofline_step']['frac'],3))
print(d['kernel_ms_per_step'])
for k,v in d.get('configs',{}).items():
    print(k, 'step', round(v['ms_per_step'],3), 'e2e', round(v['e2e']['ms_per_step'],3), 'parity', v.get('parity'), 'frac', round(v['roofline_step']['frac'],3), 'Gs/s', round(v['value']/1e9,1), v['kernel_ms_per_step'])
print(d.get('cpu_baseline'))
PY
tail -3 $out/bench_r02k.err
