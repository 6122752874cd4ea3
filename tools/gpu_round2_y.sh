#!/bin/bash
# evaluation levels + Dirichlet (separable boundary data) parity, level timing, streaming-hint A/B of the fused LSRK stage
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_apply_gpu.py tests/test_pipeline_gpu.py tests/test_cpp_driver.py -x -q -m gpu > gpurun_out/y_tests.log 2>&1; tail -3 gpurun_out/y_tests.log
HD_LIBHDGPU=$PWD/hyperdeal_b200/lib/variants/libhdgpu_fusedcs.so timeout 900 python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/y_tests_cs.log 2>&1; tail -2 gpurun_out/y_tests_cs.log
ZOO=levels timeout 600 python tools/kernel_zoo.py > gpurun_out/y_levels.txt 2>&1; cat gpurun_out/y_levels.txt
ZOO=dirichlet timeout 600 python tools/kernel_zoo.py > gpurun_out/y_dirichlet.txt 2>&1; cat gpurun_out/y_dirichlet.txt
for r in 1 2 3; do
  ZOO=lsrk timeout 300 python tools/kernel_zoo.py 2>&1 | grep "fused rk45" | sed 's/^/default  /'
  HD_LIBHDGPU=$PWD/hyperdeal_b200/lib/variants/libhdgpu_fusedcs.so ZOO=lsrk timeout 300 python tools/kernel_zoo.py 2>&1 | grep "fused rk45" | sed 's/^/fusedcs  /'
done > gpurun_out/y_fused_ab.txt 2>&1; cat gpurun_out/y_fused_ab.txt
HD_LIBHDGPU=$PWD/hyperdeal_b200/lib/variants/libhdgpu_fusedcs.so ZOO=lsrk timeout 600 ncu --set full --clock-control none -k regex:k_rounds_3d3v_k3 -s 15 -c 1 -o gpurun_out/r02y_rounds_fused_cs -f python tools/kernel_zoo.py > gpurun_out/y_ncu.log 2>&1
