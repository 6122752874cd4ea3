#!/bin/bash
# 4-GPU: peer transport A/B of the one-process-per-GPU route (torch symmetric memory vs plain allocations + CUDA IPC), parity first
mkdir -p gpurun_out
HD_PEER_TRANSPORT=ipc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 tests/mgpu_check.py > gpurun_out/z4_mgpu_ipc.log 2>&1; grep "MGPU" gpurun_out/z4_mgpu_ipc.log | cut -c1-120; tail -2 gpurun_out/z4_mgpu_ipc.log | cut -c1-200
for t in symm ipc symm ipc; do
  HD_PEER_TRANSPORT=$t timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --sustain 0 --no-cpu --no-e2e > gpurun_out/z4_bench_$t.json 2> gpurun_out/z4_bench_$t.err
  python - <<PY
import json
for l in open('gpurun_out/z4_bench_$t.json'):
    if l.startswith('{'):
        d = json.loads(l); print('$t', 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'parity_rel', d.get('parity_rel'))
PY
done
timeout 300 python tools/multi_timing.py 4 2>&1 | tee gpurun_out/z4_multi_timing.txt
