mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "exit $?" >> gpurun_out/kernel_bench.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
timeout 600 python bench.py --steps 20 --warmup 5 --graph --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "exit $?" >> gpurun_out/bench_graph.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "exit $?" >> gpurun_out/bench_ref.log
for f in pytest_gpu kernel_bench bench bench_graph bench_ref; do echo "== $f"; tail -n 30 gpurun_out/$f.log; done
