#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q -k "bidirectional or step or hot_path or config3 or corr" 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for v in "" ; do
python bench.py --steps 50 --warmup 5 --no-cpu-baseline $v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches')}), json.dumps({k:d['roofline'][k] for k in ('frac','avg_launch_ms','volumes_per_launch','kernel_share_of_step','traffic')}), d['e2e']['value'], d['clocks'])
"
done
