#!/bin/bash
# Single-GPU validation: the whole GPU suite (every test by name), config 2 / 3 timings through the C++ CLI, one bench line.
#   gpurun --timeout 3000 -- 'bash tools/gpu_validate.sh [tag]'     -> gpurun_out/<tag>_*
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/${1:-val}
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -n 2 ${O}_smoke.log
if [ -n "$FIRST_TESTS" ]; then timeout 900 python -m pytest $FIRST_TESTS -q -m gpu -rA > ${O}_pytest_first.log 2>&1; grep -E "passed|failed|error|^FAILED|^ERROR" ${O}_pytest_first.log | tail -8; fi
timeout 2400 python -m pytest tests -x -q -m gpu -rA > ${O}_pytest_all.log 2>&1
grep -E "passed|failed|error" ${O}_pytest_all.log | tail -3
MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 timeout 120 bin/miniVite_b200 -n 16777216 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror" | tee ${O}_cli.log
MVGPU_SCAN_TIMES=1 MVGPU_REPEAT=3 timeout 400 bin/miniVite_b200 -n 16777216 -p 2 -D 2>&1 | grep -E "TIMINGS|RESULT|SCAN_MS|rror" | tee -a ${O}_cli.log
timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > ${O}_bench.json 2> ${O}_bench.err
tail -c 1200 ${O}_bench.json; tail -3 ${O}_bench.err
