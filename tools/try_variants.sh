#!/bin/bash
# scratch: rebuild the decode kernel with different knobs on the GPU box and time it
for v in "$@"; do
  export DVDA_NVCC_EXTRA="$v"
  touch libdvd-audio_b200/csrc/mlp_decode.cu
  python libdvd-audio_b200/build.py > /dev/null 2>&1
  python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'step', round(d['ms_per_step'],3), 'mlp_decode', round(d['kernel_ms_per_step']['mlp_decode'],3))"
done
