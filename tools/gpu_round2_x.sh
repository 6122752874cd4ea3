#!/bin/bash
# round 2, evidence run on ONE GPU: full GPU test-suite, smoke, bench lines (headline, reference arm, k5f32, vp2d2v), kernel zoo,
# launch list, one ncu --set full capture per kernel family (read back with tools/ncu_summary.py).  gpurun merges at most 64 MiB
# back: the captures take one launch per kernel and no source import.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest_gpu.log 2>&1
tail -3 gpurun_out/x_pytest_gpu.log
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/x_smoke.log 2>&1
tail -2 gpurun_out/x_smoke.log
timeout 600 python bench.py > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
cut -c1-300 gpurun_out/x_bench.json; tail -3 gpurun_out/x_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/x_bench_ref.json 2>&1
timeout 300 python bench.py --workload k5f32 > gpurun_out/x_bench_k5f32.json 2> gpurun_out/x_bench_k5f32.err
timeout 300 python bench.py --workload vp2d2v > gpurun_out/x_bench_vp2d2v.json 2> gpurun_out/x_bench_vp2d2v.err
timeout 900 python tools/kernel_zoo.py > gpurun_out/x_kernel_zoo.txt 2>&1
ZOO=tg timeout 300 python tools/kernel_zoo.py >> gpurun_out/x_kernel_zoo.txt 2>&1
ZOO=dirichlet timeout 300 python tools/kernel_zoo.py >> gpurun_out/x_kernel_zoo.txt 2>&1
tail -8 gpurun_out/x_kernel_zoo.txt
timeout 300 python tools/vp_timing.py 1d1v 2d2v 2d2v_big > gpurun_out/x_vp.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 > gpurun_out/x_launches.log 2>&1
cap() { # name, kernel regex, launch-skip, command...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none -k regex:"$re" -s $skip -c 1 -o gpurun_out/r02x_$name -f "$@" > gpurun_out/x_ncu_$name.log 2>&1
}
ZOO=lsrk cap rounds 'k_rounds_3d3v_k3' 2 python tools/kernel_zoo.py
ZOO=lsrk cap rounds_fused 'k_rounds_3d3v_k3' 15 python tools/kernel_zoo.py
ZOO=lsrk cap norm 'k_norm_error_3d3v' 1 python tools/kernel_zoo.py
ZOO=lsrk cap interp 'k_interpolate4' 1 python tools/kernel_zoo.py
ZOO=lsrk cap stage_update 'k_stage_update' 1 python tools/kernel_zoo.py
cap vp2d2v 'k_vp_tile_2d2v' 3 python tools/vp_timing.py 2d2v
cap vsi 'k_velocity_space_integration' 1 python tools/vp_timing.py 2d2v
ZOO=tg cap tile_global 'k_apply_tile_global' 1 python tools/kernel_zoo.py
ZOO=apply2d cap tile2d2v 'k_apply_tile' 1 python tools/kernel_zoo.py
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'; du -sh gpurun_out
