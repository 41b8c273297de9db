mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_ref_n2.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
grep '^{' gpurun_out/bench_n2.log | cut -c1-300; tail -n 2 gpurun_out/bench_n2.log | cut -c1-300; grep '^{' gpurun_out/bench_ref_n2.log | cut -c1-200
