#!/bin/bash
# round 2, run B: stream priority of the cost-volume branch vs co-residency of the warp stage
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "step or hot_path" 2>&1 | tail -3 > gpurun_out/r2b.log
for prio in 0 -1 -3; do
  for v in "" ; do
  echo "== bench prio=$prio $v" >> gpurun_out/r2b.log
  STITCH_B200_COST_PRIORITY=$prio python bench.py --steps 50 --warmup 5 --no-cpu-baseline $v 2>> gpurun_out/r2b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches')}), json.dumps({k:d['roofline'][k] for k in ('frac','avg_launch_ms')}), d['e2e']['value'], d['clocks'])
" >> gpurun_out/r2b.log 2>&1
  done
done
python -c "
import torch
print('priority range', torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream,'priority_range') else 'n/a')
" >> gpurun_out/r2b.log 2>&1
cat gpurun_out/r2b.log
