#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "patch_embed" 2>&1 | tail -5
python -c "
import sys; sys.path.insert(0,'.')
import stitch_b200
print('debug word', hex(stitch_b200._lib.load().sb_debug_word()))"
timeout 300 python tools/kernel_bench.py 2>&1 | grep -E "patch_embed|useful|torch/cuDNN" | tail -4
