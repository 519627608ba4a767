#!/bin/bash
# Round 2, GPU call 5: phase-B loops on 32-bit shared addresses, one-device multi-rank tests (eager module loading), R-MAT.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/c5
timeout 900 python -m pytest tests/test_gpu_scan_kernels.py tests/test_gpu_reference_main.py -x -q -m gpu > ${O}_pytest.log 2>&1
tail -n 8 ${O}_pytest.log
run() {  # label, binary, options
  echo "== $1 opts=$3"
  MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 MVGPU_OPTIONS=$3 timeout 120 $2 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror"
}
{
  run ws bin/miniVite_b200 ""
  run pw bin/miniVite_b200 scan_variant=4
  run cap512 variants/cap512/bin/miniVite_b200 scan_variant=4
} > ${O}_cli.log 2>&1
cat ${O}_cli.log
MVGPU_OPTIONS=scan_variant=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_pw -s 11 -c 1 -o ${O}_scan_pw_it12 -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu12.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank_one_device.py -x -q -m gpu > ${O}_pytest_multirank.log 2>&1
tail -n 25 ${O}_pytest_multirank.log
for t in 8 16; do
  MVGPU_OPTIONS=scan_variant=4 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --compact-upload $t > ${O}_bench_cu$t.json 2> ${O}_bench_cu$t.err
  python - <<PY
import json
d = json.loads(open("${O}_bench_cu$t.json").read().strip().splitlines()[-1])
print("cu$t value %.4g e2e %.4g e2e_ms %.1f h2d_ms %.1f phase" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["phase_ms"]["h2d_of_e2e_step"]), d["phase_ms"])
PY
done
