#!/bin/bash
# first check of a build on the GPU box: parity suite, then the default bench on both fast paths
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_r02a.log 2>&1
echo "pytest rc=$?" >> $out/pytest_r02a.log
tail -15 $out/pytest_r02a.log
timeout 300 python bench.py --steps 10 --warmup 3 > $out/bench_r02a.json 2> $out/bench_r02a.err
DVDAGPU_THREE_PASS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_r02a_3p.json 2>> $out/bench_r02a.err
for c in c3 c4; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 2 --no-cpu-baseline > $out/bench_r02a_$c.json 2>> $out/bench_r02a.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r02a*.json')):
    try:
        d=json.load(open(f))
        print(f, d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.001})
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 $out/bench_r02a.err
