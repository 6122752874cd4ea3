#!/bin/bash
# round 2, GPU call J: next u tile loaded behind the main part
mkdir -p gpurun_out
echo "== parity" > gpurun_out/j_tests.log
timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "fast_kernel or auto_selects or fused_fast or two_bricks or self_exchange" >> gpurun_out/j_tests.log 2>&1
echo "rc=$?" >> gpurun_out/j_tests.log
tail -3 gpurun_out/j_tests.log
ROUNDS=2 timeout 1200 python tools/r6_ab.py rounds= roll=lib=r6_roll pipe=HD_FAST_VARIANT=pipe x0only=AB_VEL=1.0,0,0,0,0,0 > gpurun_out/j_ab.log 2>&1
tail -4 gpurun_out/j_ab.log
HD_LIBHDGPU=hyperdeal_b200/lib/variants/libhdgpu_r6_trace.so timeout 300 python tools/r6_timeline.py gpurun_out/j_timeline.txt > gpurun_out/j_timeline.log 2>&1
tail -17 gpurun_out/j_timeline.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rounds -s 3 -c 1 -f -o gpurun_out/r02j_rounds python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --sustain 0 > gpurun_out/j_ncu.log 2>&1
tail -1 gpurun_out/j_ncu.log | cut -c1-100
echo "== support kernels"
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q -k "interpolate or vector_tools or golden or lsrk" > gpurun_out/j_tests2.log 2>&1
tail -3 gpurun_out/j_tests2.log
ZOO=lsrk timeout 300 python tools/kernel_zoo.py > gpurun_out/j_zoo_lsrk.log 2>&1
cat gpurun_out/j_zoo_lsrk.log
