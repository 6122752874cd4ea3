#!/bin/bash
# multi-GPU: parity of the partitioned operator (tests/mgpu_check.py through pytest) and bench.py with its in-bench parity check
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/mg${N}_pytest.log 2>&1
tail -6 gpurun_out/mg${N}_pytest.log
for n in $(seq 1 4); do
  w=$((2**n)); if [ $w -gt $N ]; then break; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $((29600+w)) bench.py --gpus $w --steps 10 --warmup 5 --sustain 0 > gpurun_out/mg${N}_bench_n$w.json 2> gpurun_out/mg${N}_bench_n$w.err
  echo "== N=$w rc=$?"; tail -1 gpurun_out/mg${N}_bench_n$w.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','parity_rel','gpu_launches')}, d.get('e2e',{}).get('value'), d['config'].get('gpu_grid'), d['config'].get('halo'), d['config'].get('overlap'))
except Exception as e: print('parse failed',e)
"
  tail -3 gpurun_out/mg${N}_bench_n$w.err
done
