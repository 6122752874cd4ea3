#!/bin/bash
mkdir -p gpurun_out
HD_LIBHDGPU=hyperdeal_b200/lib/variants/libhdgpu_r6_trace.so timeout 300 python tools/r6_timeline.py gpurun_out/k_timeline_full.txt > gpurun_out/k_timeline.log 2>&1
tail -24 gpurun_out/k_timeline_full.txt
AB_VEL=1.0,0,0,0,0,0 HD_LIBHDGPU=hyperdeal_b200/lib/variants/libhdgpu_r6_trace.so timeout 300 python tools/r6_timeline.py gpurun_out/k_timeline_x0.txt >> gpurun_out/k_timeline.log 2>&1
tail -24 gpurun_out/k_timeline_x0.txt
