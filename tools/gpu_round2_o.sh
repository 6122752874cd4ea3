#!/bin/bash
mkdir -p gpurun_out
ROUNDS=2 timeout 600 python tools/r6_ab.py rounds= skipf1=lib=r6_skipf1 nopf=HD_R6_PREFETCH=0 pf45=HD_R6_PREFETCH=48 pf345=HD_R6_PREFETCH=56 x02345=AB_VEL=1.0,0,-0.05,0.1,-0.15,0.5 pipe=HD_FAST_VARIANT=pipe > gpurun_out/o_ab.log 2>&1
tail -7 gpurun_out/o_ab.log
