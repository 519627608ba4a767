#!/bin/bash
# Round 2, GPU call 7: k_scan_pq (ring of hard vertices) parity + A/B, two-ended upload.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/c7
timeout 1200 python -m pytest tests/test_gpu_scan_kernels.py -x -q -m gpu -k "5" > ${O}_pytest.log 2>&1
tail -n 8 ${O}_pytest.log
run() {  # label, binary, options
  echo "== $1 opts=$3"
  MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 MVGPU_OPTIONS=$3 timeout 120 $2 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror"
}
{
  run pw bin/miniVite_b200 ""
  run pq bin/miniVite_b200 scan_variant=5
  for v in q48 res24 qe8; do run $v variants/$v/bin/miniVite_b200 scan_variant=5; done
  run res24_pw variants/res24/bin/miniVite_b200 ""
} > ${O}_cli.log 2>&1
cat ${O}_cli.log
MVGPU_OPTIONS=scan_variant=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_pq -s 10 -c 1 -o ${O}_scan_pq_it12 -f bin/miniVite_b200 -n 16777216 -D > ${O}_ncu12.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "upload" > ${O}_pytest_upload.log 2>&1
tail -n 5 ${O}_pytest_upload.log
for mode in 1 2; do
  timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --upload-mode $mode > ${O}_bench_um$mode.json 2> ${O}_bench_um$mode.err
  python - <<PY
import json
d = json.loads(open("${O}_bench_um$mode.json").read().strip().splitlines()[-1])
print("upload mode $mode value %.4g e2e %.4g e2e_ms %.1f h2d_ms %.1f h2d_bytes %d" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["phase_ms"]["h2d_of_e2e_step"], d["e2e"]["h2d_bytes_per_step"]))
PY
done
