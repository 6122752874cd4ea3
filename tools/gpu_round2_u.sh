#!/bin/bash
# 2-GPU validation of the hd_multi_* entry points, the C++ driver on PartitionX x PartitionV GPUs, the partitioned fused LSRK
# step of the Python route, then VP stage timing
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/u_gpus.txt
timeout 900 python -m pytest tests/test_multi_capi_gpu.py tests/test_multigpu_gpu.py -x -q -m gpu > gpurun_out/u_multi.log 2>&1; echo "multi rc=$?" >> gpurun_out/u_multi.log
timeout 900 python -m pytest tests/test_cpp_driver.py tests/test_zz_vp_device_gpu.py -x -q -m gpu > gpurun_out/u_cpp.log 2>&1; echo "cpp rc=$?" >> gpurun_out/u_cpp.log
timeout 600 python tools/vp_timing.py > gpurun_out/u_vp.log 2>&1; echo "vp rc=$?" >> gpurun_out/u_vp.log
tail -5 gpurun_out/u_multi.log gpurun_out/u_cpp.log; cat gpurun_out/u_vp.log
