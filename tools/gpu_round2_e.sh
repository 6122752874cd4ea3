#!/bin/bash
# round 2, GPU call E: per-warp round-0 -> round-1 barriers; L2-only trace loads
mkdir -p gpurun_out
echo "== parity" > gpurun_out/e_tests.log
timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "fast_kernel or auto_selects or fused_fast or two_bricks or self_exchange" >> gpurun_out/e_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e_tests.log
tail -3 gpurun_out/e_tests.log
ROUNDS=2 timeout 1200 python tools/r6_ab.py rounds=HD_ROW_TILE=0,0,0,0,0 pipe=HD_FAST_VARIANT=pipe unroll=lib=r6_unroll,HD_ROW_TILE=0,0,0,0,0 cg=lib=r6_cg,HD_ROW_TILE=0,0,0,0,0 cg_unroll=lib=r6_cg_unroll,HD_ROW_TILE=0,0,0,0,0 \
  x0only=AB_VEL=1.0,0,0,0,0,0 x045=AB_VEL=1.0,0,0,0,-0.15,0.5,HD_ROW_TILE=0,0,0,0,0 x023=AB_VEL=1.0,0,-0.05,0.1,0,0,HD_ROW_TILE=0,0,0,0,0 > gpurun_out/e_ab.log 2>&1
tail -8 gpurun_out/e_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rounds -s 3 -c 1 -f -o gpurun_out/r02e_rounds python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --sustain 0 > gpurun_out/e_ncu.log 2>&1
tail -1 gpurun_out/e_ncu.log | cut -c1-100
