#!/bin/bash
# Round 2, GPU call 9: per-candidate gain bound in k_scan_pq -- parity (single rank, ranks on one device) + timing.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/c9
timeout 1500 python -m pytest tests/test_gpu_scan_kernels.py tests/test_gpu_multirank_one_device.py tests/test_gpu_rgg.py -x -q -m gpu > ${O}_pytest.log 2>&1
tail -n 6 ${O}_pytest.log
MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 timeout 120 bin/miniVite_b200 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror" | tee ${O}_cli.log
MVGPU_REPEAT=3 timeout 300 bin/miniVite_b200 -n 16777216 -p 2 2>&1 | grep -E "TIMINGS|RESULT|rror" | tee -a ${O}_cli.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_pq -s 10 -c 1 -o ${O}_scan_pq_it12 -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu12.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches.csv bin/miniVite_b200 -n 16777216 -D > /dev/null 2>&1
