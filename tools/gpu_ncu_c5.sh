#!/bin/bash
out=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_c5.csv \
    python bench.py --config c5 --steps 1 --warmup 1 --no-sub-configs --no-cpu-baseline > $out/ncu_launch_c5.log 2>&1
ls -la $out/launches_c5.csv
