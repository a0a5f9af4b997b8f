#!/bin/bash
# the sharded title set on two GPUs, both arms
out=gpurun_out
mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_r02c_n2.json 2> $out/bench_r02c_n2.err
echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $out/bench_r02c_n2_ref.json 2>> $out/bench_r02c_n2.err
echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02c_n2.json'))
print('n2', d['ms_per_step'], d['value']/1e9, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value']/1e9, 'floor', d['e2e']['copy_floor_ms_per_step'], 'parity', d['parity'])
print(d['config']); print(d.get('single_gpu'))
r=json.load(open('gpurun_out/bench_r02c_n2_ref.json')); print('ref', r['value']/1e6, r['ms_per_step'], r['cpu_baseline']['cores'])
PY
grep -v "^$" $out/bench_r02c_n2.err | tail -12
