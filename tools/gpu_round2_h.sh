#!/bin/bash
# round 2, GPU call H: lean producer warp + separate prefetch warp
mkdir -p gpurun_out
echo "== parity" > gpurun_out/h_tests.log
timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "fast_kernel or auto_selects or fused_fast or two_bricks or self_exchange" >> gpurun_out/h_tests.log 2>&1
echo "rc=$?" >> gpurun_out/h_tests.log
tail -3 gpurun_out/h_tests.log
ROUNDS=2 timeout 1200 python tools/r6_ab.py rounds_lex=HD_ROW_TILE=0,0,0,0,0 rounds= pipe=HD_FAST_VARIANT=pipe unroll=lib=r6_unroll,HD_ROW_TILE=0,0,0,0,0 nopf=HD_R6_PREFETCH=0,HD_ROW_TILE=0,0,0,0,0 \
  t22222=HD_ROW_TILE=0,2,2,2,2 stream=HD_L2_HINTS=4,HD_ROW_TILE=0,0,0,0,0 x0only=AB_VEL=1.0,0,0,0,0,0 > gpurun_out/h_ab.log 2>&1
tail -8 gpurun_out/h_ab.log
HD_LIBHDGPU=hyperdeal_b200/lib/variants/libhdgpu_r6_trace.so HD_ROW_TILE=0,0,0,0,0 timeout 300 python tools/r6_timeline.py gpurun_out/h_timeline.txt > gpurun_out/h_timeline.log 2>&1
tail -17 gpurun_out/h_timeline.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rounds -s 3 -c 1 -f -o gpurun_out/r02h_rounds python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --sustain 0 > gpurun_out/h_ncu.log 2>&1
tail -1 gpurun_out/h_ncu.log | cut -c1-100
