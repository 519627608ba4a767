#!/bin/bash
# Round 2, GPU call 10: float-build entry points (tests), config 3 (-p 2) across the three scan kernels.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/c10
timeout 900 python -m pytest tests/test_gpu_32bit.py tests/test_gpu_scan_kernels.py -x -q -m gpu -k "32 or float or mixed or golden_cases" > ${O}_pytest.log 2>&1
tail -n 12 ${O}_pytest.log
for v in 5 4 3; do
  echo "== config 3 (-n 16777216 -p 2) scan_variant=$v"
  MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 MVGPU_OPTIONS=scan_variant=$v timeout 400 bin/miniVite_b200 -n 16777216 -p 2 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror"
done > ${O}_cli.log 2>&1
cat ${O}_cli.log
MVGPU_OPTIONS=scan_variant=5 timeout 600 ncu --set full --clock-control none -k regex:k_scan_pq -s 10 -c 1 -o ${O}_scan_pq_cfg3 -f bin/miniVite_b200 -n 16777216 -p 2 > ${O}_ncu.log 2>&1
