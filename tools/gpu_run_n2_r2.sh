#!/bin/bash
cd "$(dirname "$0")/.."
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.log 2>&1
grep "^{" gpurun_out/r2_bench_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['e2e']['value'], d['cpu_baseline'], d['clocks'])
" || tail -20 gpurun_out/r2_bench_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | grep "^{" | cut -c1-400
