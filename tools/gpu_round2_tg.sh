#!/bin/bash
# degree-5 kernel with the partial sums in shared memory: thread-count variants, parity, bench line
mkdir -p gpurun_out
(for t in 448 512 672; do HD_TG_SP_THREADS=$t ZOO=tg timeout 120 python tools/kernel_zoo.py 2>&1 | grep "global-memory" | sed "s/^/smem partial sums, $t threads  /"; done
 HD_TG_SMEM_PARTIALS=0 ZOO=tg timeout 120 python tools/kernel_zoo.py 2>&1 | grep "global-memory" | sed "s/^/partial sums in dst             /") > gpurun_out/tg_sp_ab2.txt 2>&1; cat gpurun_out/tg_sp_ab2.txt
timeout 600 python -m pytest tests/test_tile_gpu.py tests/test_zz_vp_device_gpu.py tests/test_apply_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu -k "tile or k5 or global or dirichlet_on or levels or degree5 or float" > gpurun_out/tg_tests.log 2>&1; tail -3 gpurun_out/tg_tests.log
timeout 200 python bench.py --workload k5f32 > gpurun_out/tg_bench_k5f32.json 2> gpurun_out/tg_bench_k5f32.err; cut -c1-200 gpurun_out/tg_bench_k5f32.json
