#!/bin/bash
out=gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"k_mlp_filter_out|k_mlp_entropy|k_mlp_au_parse" -s 6 -c 3 -o $out/prof_c3 \
    python bench.py --config c3 --seconds 200 --steps 1 --warmup 1 --no-sub-configs --no-cpu-baseline > $out/ncu_c3.log 2>&1
ls -la $out/prof_c3.ncu-rep; tail -2 $out/ncu_c3.log | cut -c1-300
