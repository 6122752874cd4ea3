#!/bin/bash
# round 2, GPU call N: rolled tasks by default, compact waits
mkdir -p gpurun_out
echo "== parity" > gpurun_out/n_tests.log
timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "fast_kernel or auto_selects or fused_fast or two_bricks or self_exchange" >> gpurun_out/n_tests.log 2>&1
echo "rc=$?" >> gpurun_out/n_tests.log
tail -3 gpurun_out/n_tests.log
ROUNDS=3 timeout 1200 python tools/r6_ab.py rounds= unroll=lib=r6_unroll pipe=HD_FAST_VARIANT=pipe x0only=AB_VEL=1.0,0,0,0,0,0 > gpurun_out/n_ab.log 2>&1
tail -4 gpurun_out/n_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rounds -s 3 -c 1 -f -o gpurun_out/r02n_rounds python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --sustain 0 > gpurun_out/n_ncu.log 2>&1
tail -1 gpurun_out/n_ncu.log | cut -c1-100
HD_LIBHDGPU=hyperdeal_b200/lib/variants/libhdgpu_r6_trace.so timeout 300 python tools/r6_timeline.py gpurun_out/n_timeline.txt > gpurun_out/n_timeline.log 2>&1
tail -24 gpurun_out/n_timeline.txt
