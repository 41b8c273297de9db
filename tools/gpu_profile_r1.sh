# Round-1 measurement set: GPU tests, kernel microbench, bench (graph / eager), ncu launch list of the
# eager bench, ncu --set full of every hot kernel. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "exit $?" >> gpurun_out/kernel_bench.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1; echo "exit $?" >> gpurun_out/bench_eager.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "exit $?" >> gpurun_out/bench_ref.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 python tools/profile_targets.py > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"corr_umma|corr_lookup_r4|flow_warp|homo_warp|range_splat|feat_to_tokens|morph_open|attn_v_umma|softmax_rows|upsample_flow|ccl_flow|ccl_norm_tokens|tps_warp|tps_kornia" --profile-from-start off -c 40 -o gpurun_out/prof_r1 -f python tools/profile_targets.py > gpurun_out/ncu_full.log 2>&1
for f in pytest_gpu kernel_bench bench bench_eager bench_ref; do echo "== $f"; tail -n 22 gpurun_out/$f.log | cut -c1-400; done
tail -n 3 gpurun_out/ncu_full.log
