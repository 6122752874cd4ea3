#!/bin/bash
# VP: fused stages (parity + golden), CTA-shape variants of the 2D2V tile kernel, ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vp_gpu.py tests/test_zz_vp_device_gpu.py -x -q -m gpu > gpurun_out/w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/w_tests.log
for v in 0 1 2 3 4; do echo "== HD_VP_TILE_VARIANT=$v"; HD_VP_TILE_VARIANT=$v timeout 300 python tools/vp_timing.py 2d2v 2>&1 | grep "operator\|fused"; done > gpurun_out/w_variants.log 2>&1
timeout 600 python tools/vp_timing.py 1d1v 2d2v 2d2v_big > gpurun_out/w_vp.log 2>&1; echo "vp rc=$?" >> gpurun_out/w_vp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vp_tile_2d2v -c 2 -o gpurun_out/r02w_vp_tile -f python tools/vp_timing.py 2d2v > gpurun_out/w_ncu.log 2>&1
tail -n 8 gpurun_out/w_tests.log; cat gpurun_out/w_variants.log gpurun_out/w_vp.log
