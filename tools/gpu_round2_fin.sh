#!/bin/bash
# end-of-round check on ONE GPU: full GPU suite, smoke, bench lines; then env-only A/B of the fused LSRK stage
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/fin_pytest_gpu.log 2>&1; tail -3 gpurun_out/fin_pytest_gpu.log
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/fin_smoke.log 2>&1; tail -2 gpurun_out/fin_smoke.log
timeout 600 python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; cut -c1-200 gpurun_out/fin_bench.json; tail -2 gpurun_out/fin_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_bench_ref.json 2>&1; cut -c1-160 gpurun_out/fin_bench_ref.json
for v in "default" "HD_R6_PREFETCH=62" "HD_R6_PREFETCH=0" "HD_ROW_TILE=8,4,2,2,8" "HD_ROW_TILE=8,2,2,2,4" "HD_ROW_TILE=4,4,2,2,8"; do
  if [ "$v" = "default" ]; then ZOO=lsrk timeout 200 python tools/kernel_zoo.py 2>&1 | grep "again\|fused rk45" | sed "s/^/$v  /"; else env $v ZOO=lsrk timeout 200 python tools/kernel_zoo.py 2>&1 | grep "again\|fused rk45" | sed "s/^/$v  /"; fi
done > gpurun_out/fin_fused_env_ab.txt 2>&1; cat gpurun_out/fin_fused_env_ab.txt
