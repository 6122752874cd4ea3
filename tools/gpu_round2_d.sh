#!/bin/bash
# round 2, GPU call D: what makes direction 5 expensive? (address aliasing experiments)
mkdir -p gpurun_out
ROUNDS=2 timeout 1200 python tools/r6_ab.py full=HD_ROW_TILE=0,0,0,0,0 x05=AB_VEL=1.0,0,0,0,0,0.5,HD_ROW_TILE=0,0,0,0,0 x05neg=AB_VEL=1.0,0,0,0,0,-0.5,HD_ROW_TILE=0,0,0,0,0 \
   c5fast=HD_ROW_TILE=1,1,1,1,0 c5fast2=HD_ROW_TILE=2,1,1,1,0 > gpurun_out/d_ab1.log 2>&1
tail -5 gpurun_out/d_ab1.log
AB_CELLS=8,8,8,8,7,8 ROUNDS=2 timeout 600 python tools/r6_ab.py full7=HD_ROW_TILE=0,0,0,0,0 x01234_7=AB_VEL=1.0,0.15,-0.05,0.1,-0.15,0,HD_ROW_TILE=0,0,0,0,0 pipe7=HD_FAST_VARIANT=pipe > gpurun_out/d_ab2.log 2>&1
tail -3 gpurun_out/d_ab2.log
AB_CELLS=8,8,8,7,8,8 ROUNDS=1 timeout 600 python tools/r6_ab.py full_7b=HD_ROW_TILE=0,0,0,0,0 x01234_7b=AB_VEL=1.0,0.15,-0.05,0.1,-0.15,0,HD_ROW_TILE=0,0,0,0,0 > gpurun_out/d_ab3.log 2>&1
tail -2 gpurun_out/d_ab3.log
