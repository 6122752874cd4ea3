#!/bin/bash
# 4-GPU verification: hd_multi_* (C ABI) and the C++ driver on 2 and 4 GPUs, one bench line at N = 4 with NVLink counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_capi_gpu.py tests/test_cpp_driver.py -x -q -m gpu -k "multi" > gpurun_out/z_tests.log 2>&1; tail -3 gpurun_out/z_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --sustain 0 --no-cpu > gpurun_out/z_bench_n4.json 2> gpurun_out/z_bench_n4.err
grep "^{" gpurun_out/z_bench_n4.json | cut -c1-400; tail -2 gpurun_out/z_bench_n4.err
python - <<'PY'
import json
for l in open('gpurun_out/z_bench_n4.json'):
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'parity_rel', d.get('parity_rel'), 'nvlink', d.get('nvlink'), 'halo', d['config'].get('halo_bytes_sent_per_gpu_per_step'))
PY
