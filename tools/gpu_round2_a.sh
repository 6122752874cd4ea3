#!/bin/bash
# round 2, GPU call A: first run of the three-round kernel (parity, A/B against the two-role kernel, ncu), k=5 FP32 timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== parity: fast kernels" > gpurun_out/a_tests.log
timeout 900 python -m pytest tests/test_apply_gpu.py -m gpu -x -q -k "fast_kernel or auto_selects or apply_host_pipelined" >> gpurun_out/a_tests.log 2>&1
echo "rc=$?" >> gpurun_out/a_tests.log
echo "== parity: fused LSRK, ghost bricks, fused halo" >> gpurun_out/a_tests.log
timeout 1200 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q -k "fused_fast or two_bricks or self_exchange" >> gpurun_out/a_tests.log 2>&1
echo "rc=$?" >> gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
echo "== A/B"
timeout 900 python tools/r6_ab.py > gpurun_out/a_ab.log 2>&1
tail -8 gpurun_out/a_ab.log
echo "== ncu rounds kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rounds -s 3 -c 1 -f -o gpurun_out/r02a_rounds python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/a_ncu.log 2>&1
tail -3 gpurun_out/a_ncu.log
echo "== zoo (tile-global k5 f32, lsrk)"
ZOO=tg timeout 300 python tools/kernel_zoo.py > gpurun_out/a_zoo_tg.log 2>&1
ZOO=lsrk timeout 300 python tools/kernel_zoo.py > gpurun_out/a_zoo_lsrk.log 2>&1
cat gpurun_out/a_zoo_tg.log gpurun_out/a_zoo_lsrk.log | tail -12
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu > gpurun_out/a_bench.log 2>&1
tail -2 gpurun_out/a_bench.log
