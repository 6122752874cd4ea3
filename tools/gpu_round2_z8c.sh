#!/bin/bash
# 8-GPU box: bench line at N = 8 after moving the clock-sampler set-up (nvmlInit) in front of the barrier
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --sustain 0 --no-cpu > gpurun_out/z8c_bench_n8.json 2> gpurun_out/z8c_bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/z8c_bench_n8.json'):
    if l.startswith('{'):
        d = json.loads(l); print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'parity_rel', d.get('parity_rel'), 'clocks', d.get('clocks'), 'e2e', d.get('e2e', {}).get('value'))
PY
tail -2 gpurun_out/z8c_bench_n8.err | cut -c1-200
