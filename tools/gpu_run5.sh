mkdir -p gpurun_out
timeout 600 python tools/l2_granularity_exp.py > gpurun_out/l2_exp.log 2>&1; echo "exit $?" >> gpurun_out/l2_exp.log
cat gpurun_out/l2_exp.log
