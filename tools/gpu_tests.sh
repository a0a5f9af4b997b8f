#!/bin/bash
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_r02h.log 2>&1
echo "pytest rc=$?" >> $out/pytest_r02h.log
grep -E "passed|failed|FAILED|Error|assert " $out/pytest_r02h.log | tail -30
