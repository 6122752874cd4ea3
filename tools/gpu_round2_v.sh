#!/bin/bash
# VP register-tile kernels: parity (oracle) + timing + ncu of the 3D3V kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vp_gpu.py tests/test_zz_vp_device_gpu.py -x -q -m gpu > gpurun_out/v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/v_tests.log
timeout 600 python tools/vp_timing.py 3d3v 2d2v > gpurun_out/v_vp.log 2>&1; echo "vp rc=$?" >> gpurun_out/v_vp.log
timeout 600 ncu --set full --clock-control none -k regex:k_vp_tile_3d3v -s 3 -c 1 -o gpurun_out/r02v_vp3d3v -f python tools/vp_timing.py 3d3v > gpurun_out/v_ncu.log 2>&1
tail -n 15 gpurun_out/v_tests.log; cat gpurun_out/v_vp.log
