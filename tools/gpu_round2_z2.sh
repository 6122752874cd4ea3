#!/bin/bash
# 2-GPU: throughput of the hd_multi_* route, NVLink counters on the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_capi_gpu.py tests/test_cpp_driver.py -x -q -m gpu -k multi > gpurun_out/z2_tests.log 2>&1; tail -3 gpurun_out/z2_tests.log
timeout 600 python tools/multi_timing.py 1 2 > gpurun_out/z2_multi_timing.txt 2>&1; cat gpurun_out/z2_multi_timing.txt
HD_MULTI_FUSED=0 timeout 600 python tools/multi_timing.py 2 > gpurun_out/z2_multi_timing_pack.txt 2>&1; cat gpurun_out/z2_multi_timing_pack.txt
