#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "warp or homo or adapter or config1 or hot_path or step or composite or tps or smoke or range" 2>&1 | tail -15
python -c "
import sys; sys.path.insert(0,'.')
import stitch_b200
print('debug word', hex(stitch_b200._lib.load().sb_debug_word()))"
SB_BENCH_PATCH_EMBED=0 timeout 300 python tools/kernel_bench.py 2>&1 | grep -E "flow_warp|homo_warp|range_map"
