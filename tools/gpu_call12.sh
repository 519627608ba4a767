#!/bin/bash
# Round 2, GPU call 12: final single-GPU validation: whole suite, config 2 / 3 timings, ncu launch list + scan capture, bench line.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/c12
timeout 2400 python -m pytest tests -x -q -m gpu -rA > ${O}_pytest_all.log 2>&1
grep -E "passed|failed|error" ${O}_pytest_all.log | tail -3
MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 timeout 120 bin/miniVite_b200 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror" | tee ${O}_cli.log
MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 MVGPU_OPTIONS=scan_variant=5 timeout 120 bin/miniVite_b200 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror" | tee -a ${O}_cli.log
MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 timeout 400 bin/miniVite_b200 -n 16777216 -p 2 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror" | tee -a ${O}_cli.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > /dev/null 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
tail -c 2500 ${O}_bench.json; tail -3 ${O}_bench.err
