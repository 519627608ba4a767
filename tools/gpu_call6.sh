#!/bin/bash
# Round 2, GPU call 6: full single-GPU test suite on the new defaults, BFS level-byte A/B, bench line.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/c6
run() {  # label, binary, options
  echo "== $1 opts=$3"
  MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 MVGPU_OPTIONS=$3 timeout 120 $2 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror"
}
{
  run default bin/miniVite_b200 ""
  run nolvl8 variants/nolvl8/bin/miniVite_b200 ""
  run r256 bin/miniVite_b200 region_size=256
  run r1024 bin/miniVite_b200 region_size=1024
} > ${O}_cli.log 2>&1
cat ${O}_cli.log
timeout 2400 python -m pytest tests -x -q -m gpu > ${O}_pytest_all.log 2>&1
tail -n 12 ${O}_pytest_all.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msbfs -c 1 -o ${O}_msbfs -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu_bfs.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err
tail -c 1500 ${O}_bench.json; tail -3 ${O}_bench.err
