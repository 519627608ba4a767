#!/bin/bash
# 2-GPU call: gpurun --gpus 2 -- 'bash tools/gpu_multi2.sh'
# multi-GPU parity suite (NCCL + CUDA IPC between processes on different devices), patched reference main on 2 ranks,
# bench N=2 with the parity gate, in both per-iteration communication modes.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/m2
timeout 2400 python -m pytest tests -q -m gpu -k "not eight_gpus" > ${O}_pytest.log 2>&1     # the WHOLE gpu suite, 2 GPUs visible
tail -n 15 ${O}_pytest.log
for mode in 1 0; do
  MVGPU_OPTIONS=comm_mode=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > ${O}_bench_mode$mode.json 2> ${O}_bench_mode$mode.err
  python - <<PY
import json
try:
    d = json.loads(open("${O}_bench_mode$mode.json").read().strip().splitlines()[-1])
    print("comm_mode=$mode value %.4g ms %.2f e2e %.4g parity" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["parity"].get("trace_match"), d["phase_ms"], d.get("nvlink"))
except Exception as ex:
    print("comm_mode=$mode unreadable", ex)
PY
  tail -2 ${O}_bench_mode$mode.err
done
