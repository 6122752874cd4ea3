#!/bin/bash
# degree-5 global-memory tile kernel: CTA residency cap (L2 working set) A/B, then parity
mkdir -p gpurun_out
for v in "HD_TG_CTAS_PER_SM=0 HD_TG_THREADS=256" "HD_TG_CTAS_PER_SM=1 HD_TG_THREADS=512" "HD_TG_CTAS_PER_SM=1 HD_TG_THREADS=256" "HD_TG_CTAS_PER_SM=2 HD_TG_THREADS=256" "HD_TG_CTAS_PER_SM=2 HD_TG_THREADS=512" "HD_TG_CTAS_PER_SM=3 HD_TG_THREADS=256" "HD_TG_CTAS_PER_SM=4 HD_TG_THREADS=256"; do
  env $v ZOO=tg timeout 120 python tools/kernel_zoo.py 2>&1 | grep "global-memory" | sed "s/^/$v  /"
done > gpurun_out/tg_residency_ab.txt 2>&1; cat gpurun_out/tg_residency_ab.txt
ZOO=tg timeout 120 python tools/kernel_zoo.py 2>&1 | grep "global-memory" | sed "s/^/default  /" | tee -a gpurun_out/tg_residency_ab.txt
timeout 600 python -m pytest tests/test_tile_gpu.py tests/test_zz_vp_device_gpu.py tests/test_apply_gpu.py -x -q -m gpu -k "tile or k5 or global or dirichlet_on" > gpurun_out/tg_tests.log 2>&1; tail -3 gpurun_out/tg_tests.log
