#!/bin/bash
# 8-GPU call: gpurun --gpus 8 -- 'bash tools/gpu_multi8.sh'
# BASELINE.json configs[3] (RGG -n 67108864 on 8 GPUs) against the reference's 8-rank golden trace, then bench N=8 (and N=4)
# with the parity gate.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/m8
MV_CONFIG4_DEVICE_ONLY=1 timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "eight_gpus" > ${O}_pytest_config4.log 2>&1
tail -n 6 ${O}_pytest_config4.log
MVGPU_REPEAT=3 timeout 300 bin/miniVite_b200 -g 8 -n 67108864 -D > ${O}_config4_cli.log 2>&1
grep -E "TIMINGS|RESULT|Edges|Modularity" ${O}_config4_cli.log
for n in 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
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
MVGPU_OPTIONS=comm_mode=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 8 --steps 2 --warmup 3 --no-parity > ${O}_bench_n8_nccl.json 2> ${O}_bench_n8_nccl.err
tail -c 600 ${O}_bench_n8_nccl.json
