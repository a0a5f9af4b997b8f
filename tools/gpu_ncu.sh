#!/bin/bash
# ncu evidence: launch list (shares) and --set full of the main kernels, one decode of the bench track
out=gpurun_out
tag=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 60 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-sub-configs --no-cpu-baseline > $out/ncu_launch_$tag.log 2>&1
ncu --set full --import-source on --clock-control none \
    -k regex:"k_mlp_filter_out|k_mlp_entropy|k_mlp_au_parse|k_mlp_resolve|k_checkdata|k_es_gather" -s 12 -c 6 -o $out/prof_$tag \
    python bench.py --steps 1 --warmup 1 --no-sub-configs --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
ls -la $out/prof_$tag.ncu-rep $out/launches_$tag.csv
tail -3 $out/ncu_full_$tag.log
