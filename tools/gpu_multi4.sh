#!/bin/bash
# 4-GPU call: gpurun --gpus 4 -- 'bash tools/gpu_multi4.sh' : bench N=2 and N=4 (weak scaling, parity gate on).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/m4
for n in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
    bench.py --gpus $n --steps 5 --warmup 3 > ${O}_bench_n$n.json 2> ${O}_bench_n$n.err
  python - <<PY
import json
try:
    d = json.loads(open("${O}_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.4g ms %.2f e2e %.4g parity" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["parity"].get("trace_match"), d["phase_ms"], d.get("nvlink"))
except Exception as ex:
    print("N=$n unreadable", ex)
PY
  tail -2 ${O}_bench_n$n.err
done
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "not eight_gpus" -rA > ${O}_pytest_multi.log 2>&1
grep -E "passed|failed" ${O}_pytest_multi.log | tail -2
