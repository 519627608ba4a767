#!/bin/bash
# Round 2, GPU call 3: gen-3 scan v2 (TMA rings for rows+tails, 16-byte phase A, FIRST iteration), BFS done-bitmap,
# patched reference main, new bench.py parity gate.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/c4
timeout 1200 python -m pytest tests/test_gpu_scan_kernels.py tests/test_gpu_reference_main.py tests/test_gpu_parity.py -x -q -m gpu -k "not full_size" > ${O}_pytest.log 2>&1
tail -n 15 ${O}_pytest.log
run() {  # label, binary, options
  echo "== $1 opts=$3"
  MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 MVGPU_OPTIONS=$3 timeout 120 $2 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror"
}
{
  run ws bin/miniVite_b200 ""
  run pw bin/miniVite_b200 scan_variant=4
  run pw_nofirst bin/miniVite_b200 scan_variant=4,first_iter=0
  run pw_r256 bin/miniVite_b200 scan_variant=4,region_size=256
  for v in res40 cap512 cap384 w4; do run $v variants/$v/bin/miniVite_b200 scan_variant=4; done
  for v in bfsdone1 bfsdone2; do run $v variants/$v/bin/miniVite_b200 scan_variant=4; done
} > ${O}_cli.log 2>&1
cat ${O}_cli.log
MVGPU_OPTIONS=scan_variant=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches.csv bin/miniVite_b200 -n 16777216 -D > /dev/null 2>&1
MVGPU_OPTIONS=scan_variant=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_pw -s 11 -c 1 -o ${O}_scan_pw_it12 -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu12.log 2>&1
MVGPU_OPTIONS=scan_variant=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_pw -c 2 -o ${O}_scan_pw_it1 -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msbfs -c 1 -o ${O}_msbfs -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu_bfs.log 2>&1
MVGPU_OPTIONS=scan_variant=4 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err
cat ${O}_bench.json; tail -3 ${O}_bench.err
ls -la gpurun_out | grep c4_
timeout 1200 python -m pytest tests/test_gpu_multirank_one_device.py -x -q -m gpu > ${O}_pytest_multirank.log 2>&1
tail -n 25 ${O}_pytest_multirank.log
