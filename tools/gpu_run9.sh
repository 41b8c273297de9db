mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_overlap.log 2>&1; echo "exit $?" >> gpurun_out/bench_overlap.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-overlap > gpurun_out/bench_nooverlap.log 2>&1; echo "exit $?" >> gpurun_out/bench_nooverlap.log
tail -n 4 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ("bench_overlap", "bench_nooverlap"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]
        d = json.loads(l)
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "roofline frac", round(d["roofline"]["frac"], 3), "launch ms", round(d["roofline"]["avg_launch_ms"], 4))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.log").read()[-1500:])
PY
