mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "exit $?" >> gpurun_out/kernel_bench.log
tail -n 5 gpurun_out/pytest_gpu.log; cat gpurun_out/kernel_bench.log
