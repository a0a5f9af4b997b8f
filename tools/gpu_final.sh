#!/bin/bash
# what the driver runs at the end of a round, in one go
out=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/final_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > $out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/final_pytest.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/final_ref.json 2> $out/final_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > $out/final_bench.json 2> $out/final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.load(open('gpurun_out/final_ref.json')); print('reference arm:', round(r['value']/1e6,1), 'M samples/s,', r['cpu_baseline']['cores'], 'cores,', round(r['ms_per_step'],1), 'ms/step')
d=json.load(open('gpurun_out/final_bench.json'))
print('headline', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,1), 'G; e2e', round(d['e2e']['ms_per_step'],2), 'ms', round(d['e2e']['value']/1e9,2), 'G; parity', d['parity'], 'frac', round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], 'launches', d['gpu_launches'], 'steps', d['steps'])
for k,v in d.get('configs',{}).items(): print(k, round(v['ms_per_step'],3), round(v['e2e']['ms_per_step'],2), v.get('parity'))
print(d['clocks'])
PY
