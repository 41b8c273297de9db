#!/bin/bash
# round 2, run A: corr register/smem trim (co-residency of the warp stage), lookup sub-batching
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "corr or step or hot_path or gma or ccl" 2>&1 | tail -5 > gpurun_out/r2a_tests.log
for v in "" "--no-overlap" "--lookup-subbatch 4" "--lookup-subbatch 8" "--lookup-subbatch 2"; do
  echo "== bench $v" >> gpurun_out/r2a_bench.log
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline $v 2>> gpurun_out/r2a_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches')}), json.dumps({k:d['roofline'][k] for k in ('frac','avg_launch_ms')}), json.dumps({k:d['e2e'].get(k) for k in ('value','pcie_gbs','h2d_peak_gbs')}), json.dumps(d['e2e'].get('images_only') and d['e2e']['images_only']['value']), d['clocks'])
" >> gpurun_out/r2a_bench.log 2>&1
done
cat gpurun_out/r2a_tests.log gpurun_out/r2a_bench.log
