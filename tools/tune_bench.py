"""bench.py under one sb_tune key: python tools/tune_bench.py <key> <v1,v2,...> (run from the repo root)."""
import sys, json, io, contextlib, runpy
sys.path.insert(0, ".")
import stitch_b200 as sb
lib = sb._lib.load()
key, vals = int(sys.argv[1]), [int(v) for v in sys.argv[2].split(",")]
for v in vals:
    lib.sb_tune(key, v)
    sys.argv = ["bench.py", "--steps", "20", "--warmup", "5", "--no-cpu-baseline"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        try:
            runpy.run_path("bench.py", run_name="__main__")
        except SystemExit:
            pass
    line = [l for l in buf.getvalue().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print("tune(%d, %d): %.0f pairs/s, %.4f ms" % (key, v, d["value"], d["ms_per_step"]), flush=True)
