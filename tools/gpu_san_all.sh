#!/bin/bash
# memcheck over the catalog-sized parity tests (the large workloads are left out: the tool is 20-50x slower)
out=gpurun_out
timeout 1100 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_parity.py \
  -k "tracks_one_by_one or whole_titleset or sync_search or tables_sized or damage or truncated or sharded or device_resident or (parts_concatenate and not large) or (pipelined and not large)" \
  -q > $out/san_all.log 2>&1
echo rc=$?
grep -n "Invalid\|at .*k_\|passed\|failed\|ERROR SUMMARY" $out/san_all.log | head -30
