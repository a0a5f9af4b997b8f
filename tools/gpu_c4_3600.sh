#!/bin/bash
# verdict item: C4 at its full length (3600 s of 24-bit 192 kHz stereo MLP) once, with parity
out=gpurun_out
timeout 1500 python bench.py --config c4 --seconds 3600 --steps 5 --warmup 3 --no-sub-configs --no-cpu-baseline > $out/bench_r02_c4_3600.json 2> $out/bench_r02_c4_3600.err
echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_c4_3600.json'))
print('c4 3600 s: step', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,1), 'Gs/s; e2e', round(d['e2e']['ms_per_step'],2), 'ms', round(d['e2e']['value']/1e9,2), 'Gs/s; parity', d['parity'], d['config'])
print(d['kernel_ms_per_step'])
PY
tail -3 $out/bench_r02_c4_3600.err; nvidia-smi --query-gpu=memory.used --format=csv
