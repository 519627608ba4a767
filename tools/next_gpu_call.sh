#!/bin/bash
# First GPU call of the next round: measures everything that was prepared after round 1's GPU budget ran out.
#   gpurun --timeout 1500 -- 'bash tools/next_gpu_call.sh'
# Part 1 uses the C++ CLI only (no Python start-up): ~6 s of GPU time per line.  Results land in gpurun_out/next_*.log.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
{
  echo "== config 2 (-n 16777216), 3 warm repeats, options one at a time"
  for o in "" degree_sort=256 degree_sort=512 degree_sort=1024 degree_sort=2048 fold_variant=1 "degree_sort=1024,fold_variant=1" region_size=256 region_size=1024; do
    echo "opts=$o"
    MVGPU_REPEAT=3 MVGPU_OPTIONS=$o timeout 120 bin/miniVite_b200 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|rror"
  done
  echo "== config 3-like irregular graph is host-generated (-p 2): default vs degree_sort"
  for o in "" degree_sort=1024; do
    echo "opts=$o"
    MVGPU_REPEAT=3 MVGPU_OPTIONS=$o timeout 300 bin/miniVite_b200 -n 16777216 -p 2 2>&1 | grep -E "TIMINGS|RESULT|rror"
  done
  echo "== compile-time variants built with tools/build_variant.sh (if any)"
  for b in variants/*/bin/miniVite_b200; do
    [ -x "$b" ] || continue
    for o in "" degree_sort=1024; do
      echo "variant=$b opts=$o"
      MVGPU_REPEAT=3 MVGPU_OPTIONS=$o timeout 120 $b -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|rror"
    done
  done
} > gpurun_out/next_cli.log 2>&1
# Part 2: parity of everything new (tests added after the last GPU run of round 1)
timeout 900 python -m pytest tests -q -m gpu -k "survey or experimental or renumbering or unit_weight_cases" > gpurun_out/next_pytest.log 2>&1
# Part 3: e2e with the AVX2 compact upload
for t in 0 16 32 64; do
  timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --compact-upload $t > gpurun_out/next_bench_cu$t.json 2> gpurun_out/next_bench_cu$t.err
done
tail -n 40 gpurun_out/next_cli.log; tail -n 5 gpurun_out/next_pytest.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/next_bench_cu*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "e2e ms %.1f" % d["e2e"]["ms_per_step"], d["phase_ms"])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
