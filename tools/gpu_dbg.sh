#!/bin/bash
out=gpurun_out
DVDAGPU_DEBUG=1 timeout 600 python -m pytest tests -m gpu -x -q > $out/dbg.log 2>&1
grep -E "attempt|buffer|passed|failed" $out/dbg.log | tail -40
