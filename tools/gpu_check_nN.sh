#!/bin/bash
# the sharded title set on N GPUs (our arm; the reference arm once, it does not depend on N much)
N=${1:-2}
out=gpurun_out
mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_r02m_n$N.json 2> $out/bench_r02m_n$N.err
echo "bench rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d=json.load(open('gpurun_out/bench_r02m_n%s.json' % n))
print('n', n, 'step ms', round(d['ms_per_step'],3), 'Gs/s', round(d['value']/1e9,1), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'Gs/s', round(d['e2e']['value']/1e9,2), 'floor', round(d['e2e']['copy_floor_ms_per_step'],2), 'parity', d['parity'])
print(d["config"]["units_per_rank"], d["config"]["sectors_per_rank"], d["config"].get("step_ms_per_rank"), d["config"].get("e2e_ms_per_rank")); print(d.get('single_gpu'))
PY
grep -v "^$\|\*\*\*\|OMP_NUM" $out/bench_r02m_n$N.err | tail -6
