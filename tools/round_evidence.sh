#!/bin/bash
# The round's GPU evidence in one go (run under gpurun, one GPU):
#   tools/round_evidence.sh <tag>        -> gpurun_out/*_<tag>.*
# 1. the numbers quoted as results: bench.py not under a profiler (default workload, then the other configurations)
# 2. the reference arm
# 3. ncu launch list of one bench run (cold, serialised: compare shares, not absolutes)
# 4. ncu --set full of the main kernels
# 5. the time line of one decode (DVDAGPU_TRACE)
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py --steps 10 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_${tag}_ref.json 2>> $out/bench_$tag.err
for c in c1 c3 c4; do
  python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_${tag}_$c.json 2>> $out/bench_$tag.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --seconds 600 --no-cpu-baseline > $out/bench_under_ncu_$tag.log 2>&1
ncu --set full --import-source on --clock-control none \
    -k regex:"k_mlp_filter_out|k_mlp_entropy|k_mlp_au_parse|k_checkdata|k_es_gather|k_sync_find" -c 6 -o $out/prof_$tag \
    python bench.py --steps 1 --warmup 0 --seconds 600 --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
python tools/trace_decode.py 600 > $out/trace_$tag.txt 2>&1
ls -la $out/*_$tag*
