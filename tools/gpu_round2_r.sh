#!/bin/bash
# round 2, GPU call R: full GPU test-suite, bench line, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest_gpu.log 2>&1
tail -5 gpurun_out/r_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
cut -c1-1500 gpurun_out/r_bench.json; tail -3 gpurun_out/r_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r_bench_ref.json 2>&1
cut -c1-600 gpurun_out/r_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 > gpurun_out/r_launches.log 2>&1
tail -2 gpurun_out/r_launches.log | cut -c1-200
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r_smoke.log 2>&1
tail -3 gpurun_out/r_smoke.log
