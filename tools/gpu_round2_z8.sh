#!/bin/bash
# 8-GPU box: hd_multi_* parity at 8 GPUs, throughput of the C++ route at 1/2/4/8, bench line at N = 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_capi_gpu.py -x -q -m gpu > gpurun_out/z8_tests.log 2>&1; tail -3 gpurun_out/z8_tests.log
timeout 900 python tools/multi_timing.py 1 2 4 8 > gpurun_out/z8_multi_timing.txt 2>&1; cat gpurun_out/z8_multi_timing.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --sustain 0 --no-cpu > gpurun_out/z8_bench_n8.json 2> gpurun_out/z8_bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/z8_bench_n8.json'):
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'parity_rel', d.get('parity_rel'), 'e2e', d.get('e2e', {}).get('value'))
PY
timeout 300 env HD_PARTITION_X=4 HD_PARTITION_V=2 hyperdeal_b200/bin/advection tests/golden/adv_2D_2D_k3.hyperrectangle_03.json > gpurun_out/z8_cpp_driver_8gpu.txt 2>&1; tail -4 gpurun_out/z8_cpp_driver_8gpu.txt
