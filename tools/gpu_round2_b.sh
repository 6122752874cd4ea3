#!/bin/bash
# round 2, GPU call B: three-round kernel with the consume-then-request trace pipeline
mkdir -p gpurun_out
echo "== parity" > gpurun_out/b_tests.log
timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "fast_kernel or auto_selects or fused_fast or two_bricks or self_exchange" >> gpurun_out/b_tests.log 2>&1
echo "rc=$?" >> gpurun_out/b_tests.log
tail -4 gpurun_out/b_tests.log
echo "== A/B"
ROUNDS=2 timeout 1200 python tools/r6_ab.py pipe=HD_FAST_VARIANT=pipe rounds=HD_FAST_VARIANT=rounds rounds_nopf=HD_R6_PREFETCH=0 rounds_stream=HD_L2_HINTS=4 \
   rounds_t22222=HD_ROW_TILE=0,2,2,2,2 rounds_t2222s=HD_ROW_TILE=0,2,2,2,2,HD_L2_HINTS=4 rounds_t4220=HD_ROW_TILE=0,4,2,2,0 rounds_lex=HD_ROW_TILE=0,0,0,0,0 \
   rounds_unroll=lib=r6_unroll \
   pipe_x0only=HD_FAST_VARIANT=pipe,AB_VEL=1.0,0,0,0,0,0 rounds_x0only=AB_VEL=1.0,0,0,0,0,0 rounds_x01=AB_VEL=1.0,0.15,0,0,0,0 > gpurun_out/b_ab.log 2>&1
tail -14 gpurun_out/b_ab.log
echo "== ncu rounds kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rounds -s 3 -c 1 -f -o gpurun_out/r02b_rounds python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --sustain 0 > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/b_ncu.log | cut -c1-300
