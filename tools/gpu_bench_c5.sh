#!/bin/bash
# c5 on one GPU (both legs) — the batched end-to-end leg
out=gpurun_out
timeout 600 python bench.py --config c5 --steps 10 --warmup 3 --no-sub-configs --no-cpu-baseline > $out/bench_r02n_c5.json 2> $out/bench_r02n_c5.err
echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02n_c5.json'))
print('c5 step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), 'parity', d['parity'], d['e2e'].get('gpu_launches'))
PY
tail -3 $out/bench_r02n_c5.err
