#!/bin/bash
# Round 2, GPU call 11: the whole single-GPU suite on the final defaults (scan_variant 6), config 2 / config 3 timings.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/c11
timeout 2400 python -m pytest tests -x -q -m gpu -rA > ${O}_pytest_all.log 2>&1
grep -E "passed|failed|error" ${O}_pytest_all.log | tail -3
MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 timeout 120 bin/miniVite_b200 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror" | tee ${O}_cli.log
MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 timeout 400 bin/miniVite_b200 -n 16777216 -p 2 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror" | tee -a ${O}_cli.log
