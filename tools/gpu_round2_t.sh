#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q -k "self_exchange or two_bricks" > gpurun_out/t_tests.log 2>&1
tail -3 gpurun_out/t_tests.log
CELLS=16,8,4,8,8,8 DIRS=1,2 SENDERS=16,32,64 timeout 300 python tools/fused_selftest.py 2>&1 | tee gpurun_out/t_selftest_defer.log
HD_R6_DEFER=0 CELLS=16,8,4,8,8,8 DIRS=1,2 SENDERS=32 timeout 300 python tools/fused_selftest.py 2>&1 | tail -3 | tee gpurun_out/t_selftest_lists.log
