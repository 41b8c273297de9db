#!/bin/bash
# Round-2 measurement set: GPU tests, kernel microbench, bench (graph / batch 64 / reference arm), ncu launch list of the
# eager bench, ncu --set full of the hot kernels. Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/kernel_bench.py > gpurun_out/r2_kernel_bench.log 2>&1; echo "exit $?" >> gpurun_out/r2_kernel_bench.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.log 2>&1; echo "exit $?" >> gpurun_out/r2_bench.log
timeout 600 python bench.py --steps 20 --warmup 5 --batch 64 --no-cpu-baseline > gpurun_out/r2_bench_b64.log 2>&1; echo "exit $?" >> gpurun_out/r2_bench_b64.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1; echo "exit $?" >> gpurun_out/r2_bench_ref.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 python tools/profile_targets.py > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"corr_umma|corr_lookup_r4|flow_warp|homo_warp|range_splat|range_finalize|feat_to_tokens|patch_embed_umma" --profile-from-start off -c 40 -o gpurun_out/prof_r2 -f python tools/profile_targets.py > gpurun_out/ncu_full.log 2>&1
timeout 600 python tools/config45_bench.py > gpurun_out/r2_configs_4_5.log 2>&1; echo "exit $?" >> gpurun_out/r2_configs_4_5.log
for f in r2_configs_4_5 r2_pytest_gpu r2_kernel_bench r2_bench r2_bench_b64 r2_bench_ref; do echo "== $f"; tail -n 40 gpurun_out/$f.log | cut -c1-600; done
tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
