mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 400 python tools/corr_diag.py > gpurun_out/corr_diag.log 2>&1; echo "corr_diag exit $?" >> gpurun_out/corr_diag.log
timeout 900 python -m pytest tests -m gpu -q -k "not corr and not hot_path" -p no:cacheprovider > gpurun_out/pytest_nocorr.log 2>&1; echo "exit $?" >> gpurun_out/pytest_nocorr.log
timeout 900 python -m pytest tests -m gpu -q -k "corr or hot_path" -p no:cacheprovider > gpurun_out/pytest_corr.log 2>&1; echo "exit $?" >> gpurun_out/pytest_corr.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
tail -5 gpurun_out/corr_diag.log gpurun_out/pytest_nocorr.log gpurun_out/pytest_corr.log gpurun_out/bench.log
