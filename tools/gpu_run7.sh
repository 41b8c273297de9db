mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "lookup or step or 1024" > gpurun_out/pytest_lookup.log 2>&1; echo "exit $?" >> gpurun_out/pytest_lookup.log
timeout 300 python tools/lookup_exp.py > gpurun_out/lookup_exp.log 2>&1; echo "exit $?" >> gpurun_out/lookup_exp.log
tail -n 8 gpurun_out/pytest_lookup.log; cat gpurun_out/lookup_exp.log
