"""Turn the ncu outputs a gpurun call brought back (gpurun_out/) into the small, tracked
summaries under profiles/:  <tag>_launches.md  (per-kernel share of the step, from the
gpu__time_duration launch list) and <tag>_ncu_full.md (+ corr_umma_traffic.json) from the
--set full capture."""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = os.path.join(ROOT, "gpurun_out")


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("sb::", "")
    return name.strip()


def launches():
    path = os.path.join(src, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r[4])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over "
                f"`python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline`\n\n"
                "Per-launch times are cold-cache and serialised; read the SHARES.\n"
                f"{len(rows)} launches captured (warm-up + timed + e2e passes), total {tot/1e3:.2f} ms.\n\n"
                "| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {us/n:.1f} | {100*us/tot:.1f}% |\n")
    print("wrote", f"{tag}_launches.md")


def full():
    rep = os.path.join(src, f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
            "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "smsp__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tc.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]

    def to_bytes(v, u):
        v = float(v)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    with open(os.path.join(OUT, f"{tag}_ncu_full.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none ({tag}), tools/profile_targets.py at B=16, 512^2\n\n")
        for r in data:
            name = short(r[col["Kernel Name"]])
            f.write(f"## `{name}`\n\n| metric | value |\n|---|---|\n")
            for w in want:
                if w in col:
                    f.write(f"| {w} | {r[col[w]]} {units[col[w]]} |\n")
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            f.write(f"| dram traffic (read+write) | {(rd+wr)/1e6:.1f} MB |\n\n")
            if name.startswith("corr_umma_kernel<1>") or name.startswith("corr_umma_kernel<1,"):
                json.dump({"kernel": name, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                           "source": f"profiles/{tag}_ncu_full.md (ncu --set full --clock-control none, tools/profile_targets.py)"}, open(os.path.join(OUT, "corr_umma_traffic.json"), "w"))
    print("wrote", f"{tag}_ncu_full.md")


os.makedirs(OUT, exist_ok=True)
launches()
full()
