#!/bin/bash
# Round 2, GPU call 2: parity of the new kernels (k_scan_pw, sub-warp BFS, narrowed sort, 16-byte fold), then A/B timings.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/c2
timeout 900 python -m pytest tests/test_gpu_scan_kernels.py tests/test_gpu_parity.py -x -q -m gpu -k "not full_size" > ${O}_pytest.log 2>&1
tail -n 15 ${O}_pytest.log
run() {  # label, binary, options
  echo "== $1 opts=$3"
  MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 MVGPU_OPTIONS=$3 timeout 120 $2 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror"
}
{
  run default bin/miniVite_b200 ""
  run pw bin/miniVite_b200 scan_variant=4
  run pw_nopol bin/miniVite_b200 scan_variant=4,cache_policy=0
  run pw_r256 bin/miniVite_b200 scan_variant=4,region_size=256
  for v in bfs1 bfs4 bfs16; do run $v variants/$v/bin/miniVite_b200 ""; done
  for v in cap512 w4; do run $v variants/$v/bin/miniVite_b200 scan_variant=4; done
} > ${O}_cli.log 2>&1
cat ${O}_cli.log
# launch list of one warm phase with the new kernel + full captures of k_scan_pw (iterations 1-2 and 12-13)
MVGPU_OPTIONS=scan_variant=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches.csv bin/miniVite_b200 -n 16777216 -D > /dev/null 2>&1
MVGPU_OPTIONS=scan_variant=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_pw -s 11 -c 2 -o ${O}_scan_pw_it12 -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu12.log 2>&1
MVGPU_OPTIONS=scan_variant=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_pw -c 2 -o ${O}_scan_pw_it1 -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msbfs -c 1 -o ${O}_msbfs -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu_bfs.log 2>&1
ls -la gpurun_out | grep c2_
