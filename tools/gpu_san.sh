#!/bin/bash
out=gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -k "tables_sized_in_advance_grow" -x -q > $out/san.log 2>&1
echo rc=$?
grep -n "Invalid\|at .*k_\|by thread\|Address\|passed\|failed" $out/san.log | head -30
