#!/bin/bash
# round 2, GPU call M: direction-1 face layer staged by TMA, two P buffers
mkdir -p gpurun_out
echo "== parity" > gpurun_out/m_tests.log
timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "fast_kernel or auto_selects or fused_fast or two_bricks or self_exchange" >> gpurun_out/m_tests.log 2>&1
echo "rc=$?" >> gpurun_out/m_tests.log
tail -3 gpurun_out/m_tests.log
ROUNDS=3 timeout 1200 python tools/r6_ab.py rounds= roll=lib=r6_roll pipe=HD_FAST_VARIANT=pipe x0only=AB_VEL=1.0,0,0,0,0,0 x02345=AB_VEL=1.0,0,-0.05,0.1,-0.15,0.5 > gpurun_out/m_ab.log 2>&1
tail -5 gpurun_out/m_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rounds -s 3 -c 1 -f -o gpurun_out/r02m_rounds python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --sustain 0 > gpurun_out/m_ncu.log 2>&1
tail -1 gpurun_out/m_ncu.log | cut -c1-100
