#!/bin/bash
mkdir -p gpurun_out
ROUNDS=2 timeout 600 python tools/r6_ab.py rounds= a1tiny=AB_VEL=1.0,1e-30,-0.05,0.1,-0.15,0.5 x02345=AB_VEL=1.0,0,-0.05,0.1,-0.15,0.5 alltiny=AB_VEL=1e-30,1e-30,-1e-30,1e-30,-1e-30,1e-30 x0only=AB_VEL=1.0,0,0,0,0,0 > gpurun_out/p_ab.log 2>&1
tail -5 gpurun_out/p_ab.log
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
