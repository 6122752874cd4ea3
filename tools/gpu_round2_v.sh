#!/bin/bash
# VP register-tile kernels: parity (oracle) + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vp_gpu.py tests/test_zz_vp_device_gpu.py -x -q -m gpu > gpurun_out/v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/v_tests.log
timeout 600 python tools/vp_timing.py 1d1v 2d2v 2d2v_big > gpurun_out/v_vp.log 2>&1; echo "vp rc=$?" >> gpurun_out/v_vp.log
tail -n 15 gpurun_out/v_tests.log; cat gpurun_out/v_vp.log
