#!/bin/bash
out=gpurun_out
DVDA_B200_DEBUG=1 timeout 600 python -m pytest tests -m gpu -x -q -k "reads_long_tracks_in_parts and mlp_param_dup" > $out/dbg.log 2>&1
grep -E "dvda\]|passed|failed" $out/dbg.log | tail -40
