#!/bin/bash
# round 2, GPU call C: three-round kernel with round 2 adding the partial sums at the end (decoupled pipeline)
mkdir -p gpurun_out
echo "== parity" > gpurun_out/c_tests.log
timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "fast_kernel or auto_selects or fused_fast or two_bricks or self_exchange" >> gpurun_out/c_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c_tests.log
tail -4 gpurun_out/c_tests.log
echo "== A/B"
ROUNDS=2 timeout 1200 python tools/r6_ab.py pipe=HD_FAST_VARIANT=pipe rounds=HD_FAST_VARIANT=rounds rounds_nopf=HD_R6_PREFETCH=0 \
   rounds_t22222=HD_ROW_TILE=0,2,2,2,2 rounds_lex=HD_ROW_TILE=0,0,0,0,0 rounds_unroll=lib=r6_unroll \
   rounds_x0only=AB_VEL=1.0,0,0,0,0,0 rounds_x012=AB_VEL=1.0,0.15,-0.05,0,0,0,HD_ROW_TILE=0,0,0,0,0 rounds_x0123=AB_VEL=1.0,0.15,-0.05,0.1,0,0,HD_ROW_TILE=0,0,0,0,0 \
   rounds_x01234=AB_VEL=1.0,0.15,-0.05,0.1,-0.15,0,HD_ROW_TILE=0,0,0,0,0 > gpurun_out/c_ab.log 2>&1
tail -11 gpurun_out/c_ab.log
echo "== ncu rounds kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rounds -s 3 -c 1 -f -o gpurun_out/r02c_rounds python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --sustain 0 > gpurun_out/c_ncu.log 2>&1
tail -2 gpurun_out/c_ncu.log | cut -c1-200
ZOO=lsrk timeout 300 python tools/kernel_zoo.py > gpurun_out/c_zoo_lsrk.log 2>&1
head -3 gpurun_out/c_zoo_lsrk.log
