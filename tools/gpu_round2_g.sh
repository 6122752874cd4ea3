#!/bin/bash
mkdir -p gpurun_out
HD_LIBHDGPU=hyperdeal_b200/lib/variants/libhdgpu_r6_trace.so timeout 300 python tools/r6_timeline.py gpurun_out/g_timeline.txt > gpurun_out/g_timeline.log 2>&1
tail -22 gpurun_out/g_timeline.txt
ROUNDS=1 timeout 600 python tools/r6_ab.py x02345=AB_VEL=1.0,0,-0.05,0.1,-0.15,0.5,HD_ROW_TILE=0,0,0,0,0 x12345=AB_VEL=0,0.15,-0.05,0.1,-0.15,0.5,HD_ROW_TILE=0,0,0,0,0 x0145=AB_VEL=1.0,0.15,0,0,-0.15,0.5,HD_ROW_TILE=0,0,0,0,0 x02345s=AB_VEL=1.0,0,-0.05,0.1,-0.15,0.5,HD_ROW_TILE=0,0,0,0,0,HD_L2_HINTS=4 > gpurun_out/g_ab.log 2>&1
tail -4 gpurun_out/g_ab.log
