#!/bin/bash
# 8-GPU box: clean bench lines at N = 8 for both peer transports (no NVML in the timed path)
mkdir -p gpurun_out
for t in symm ipc; do
  HD_PEER_TRANSPORT=$t timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --sustain 0 --no-cpu --no-e2e > gpurun_out/z8b_bench_$t.json 2> gpurun_out/z8b_bench_$t.err
  python - <<PY
import json
for l in open('gpurun_out/z8b_bench_$t.json'):
    if l.startswith('{'):
        d = json.loads(l); print('$t', 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'parity_rel', d.get('parity_rel'))
PY
done
timeout 200 python tools/multi_timing.py 8 2>&1 | tee gpurun_out/z8b_multi_timing.txt
