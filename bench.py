#!/usr/bin/env python
"""bench.py — 512x512 pairs/s through the stitching-alignment hot path
(cost volume + lookup + warp), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One *step* = one batch of 16 synthetic UDIS-D-shaped pairs per GPU through the op
list of the reference's ``train_eval_foward`` (2 x cost volume (+pyramid), 24 x
lookup, 2 x homography warp, occlusion, flow warp).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "512x512 pairs/sec (cost volume+warp)"
UNIT = "pairs/s"
BATCH_PER_GPU = 16
SIZE = 512
ITERS = 12


def workload_name(n_gpus):
    return (f"synthetic UDIS-D 512x512 pairs, batch {BATCH_PER_GPU} per GPU, cost volume (+pyramid) + "
            f"24 lookups + 2 homography warps + occlusion + flow warp on {n_gpus}xB200")


# --------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------- CPU baseline
def cpu_pairs_per_s(n_pairs, min_seconds, max_seconds=60.0):
    """The reference's CPU path for the same op list, as the oracle port (numpy BLAS for
    the contraction, OpenMP C for the gathers), all host threads. Returns (pairs/s, pairs done)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import stitch_oracle as so
    from stitch_b200.pipeline import make_pair_batch
    pb = make_pair_batch(0, n_pairs, size=SIZE, iters=ITERS)
    a = {k: v.numpy() for k, v in zip(("image1", "image2", "fmap1", "fmap2", "h_motion", "flow_ij", "flow_ji", "coords"),
                                      pb.tensors())}

    def one_pass():
        b = n_pairs
        for f1, f2 in ((a["fmap1"], a["fmap2"]), (a["fmap2"], a["fmap1"])):
            pyr = so.corr_pyramid(f1, f2, 4)
            for it in range(ITERS):
                so.encode_flow_token(pyr[0], a["coords"][it])
        src = np.tile(np.array([[0.0, 0.0], [SIZE, 0.0], [0.0, SIZE], [SIZE, SIZE]], np.float32)[None], (b, 1, 1))
        H = so.tensor_DLT(src / 8, (src + a["h_motion"]) / 8)
        M = np.array([[SIZE / 16.0, 0, SIZE / 16.0], [0, SIZE / 16.0, SIZE / 16.0], [0, 0, 1]], np.float32)
        Mi = np.linalg.inv(M).astype(np.float32)
        H_mat = (Mi @ H @ M).astype(np.float32)
        H_inv = (Mi @ np.linalg.inv(H).astype(np.float32) @ M).astype(np.float32)
        ones = np.ones_like(a["image2"])
        out_h = so.homo_transformer(np.concatenate((a["image2"], ones), 1), H_mat, (SIZE, SIZE))
        so.homo_transformer(np.concatenate((a["image1"], ones), 1), H_inv, (SIZE, SIZE))
        occ = so.compute_occlusion_wang(a["flow_ji"], True, threshold=True)
        so.warp(out_h, a["flow_ij"], mul_mask=occ, return_overlap=True)

    one_pass()  # warm-up (library loading, page faults)
    done, t0 = 0, time.perf_counter()
    while True:
        one_pass()
        done += n_pairs
        el = time.perf_counter() - t0
        if el >= min_seconds or el >= max_seconds:
            break
    return done / el, done


def cpu_model() -> str:
    """CPU model string of the host the CPU legs ran on (reported with the numbers)."""
    try:
        for l in open("/proc/cpuinfo"):
            if l.lower().startswith("model name"):
                return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path. The reference is
    pure Python and cannot travel to the GPU box (no /root/reference there), so the timed code is
    the oracle port (cpu_baseline.kind = "port"). Rank 0 only; other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sample_pairs = 2
    per_step = []
    cpu_pairs_per_s(sample_pairs, 0.0)  # warm
    for i in range(args.warmup + args.steps):
        v, _ = cpu_pairs_per_s(sample_pairs, 0.0)
        if i >= args.warmup:
            per_step.append(v)
    value = statistics.mean(per_step)
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sample_pairs / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "note": "CPU arm: each step is a bounded sample of "
                   f"{sample_pairs} pairs of the same workload on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                         "sample": f"{sample_pairs} pairs per step x {args.steps} steps, all host threads (OpenMP + BLAS)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import stitch_b200
    from stitch_b200 import _lib
    from stitch_b200.pipeline import HotPath, algorithmic_work, make_pair_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().sb_device_check(), "sb_device_check")

    B = BATCH_PER_GPU
    # rank r owns pairs [r*B, (r+1)*B) of the global list (weak scaling, no data-path collective)
    # pinned host staging buffers (--wc: write-combined, filled once by the CPU, read by the device)
    pb_host = make_pair_batch(rank * B, B, size=SIZE, iters=ITERS).map(
        (lambda t: _lib.pinned_like(t, write_combined=True)) if args.wc else (lambda t: t.pin_memory()))
    pb_dev = pb_host.map(lambda t: t.to(dev, non_blocking=True))
    hp = HotPath(size=SIZE, iters=ITERS, pyramid=True, overlap=args.overlap, eval_outputs=not args.graph)
    stream = torch.cuda.current_stream()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    out = None
    for _ in range(args.warmup):
        out = hp.step(pb_dev)       # keeps one result set alive like the timed loop (same allocator footprint)
    barrier()
    run_step = lambda: hp.step(pb_dev)
    _lib.reset_launch_count()
    hp.step(pb_dev)
    launches_per_step_eager = _lib.launch_count()
    if args.graph:
        # the step is sync-free: capture its launches once, replay as one submission
        hp.capture(pb_dev)
        run_step = hp.replay
        for _ in range(2):
            run_step()
        barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    hp.gemm_events = None if args.graph else []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        out = run_step()
    e1.record(stream)
    barrier()
    launches = _lib.launch_count()
    ms_total = e0.elapsed_time(e1)
    gemm_ms = [a.elapsed_time(b) for a, b in (hp.gemm_events or [])]
    hp.gemm_events = None
    ms_eager_total = ms_total
    if args.graph:
        # Events cannot be timed inside a captured graph, and an eager pass of the whole step is
        # host-bound (event pairs would include launch gaps). The dominant kernel is therefore
        # timed right after the timed region as back-to-back launches on the same stream, same
        # operands, each launch writing its full 1.33 GiB of outputs (>> L2).
        launches = launches_per_step_eager * args.steps
        from stitch_b200 import corr as corr_mod
        s8 = SIZE // 8
        tk1, tk2 = corr_mod.tokens_bf16(pb_dev.fmap1), corr_mod.tokens_bf16(pb_dev.fmap2)
        for _ in range(3):
            corr_mod.corr_from_tokens(tk1, tk2, 256, (s8, s8), (s8, s8), pyramid_levels=3)
        n_rep = 20
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_rep + 1)]
        evs[0].record(stream)
        for i in range(n_rep):
            corr_mod.corr_from_tokens(tk1 if i % 2 == 0 else tk2, tk2 if i % 2 == 0 else tk1, 256, (s8, s8), (s8, s8),
                                      pyramid_levels=3)
            evs[i + 1].record(stream)
        barrier()
        gemm_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_rep)]
        ms_eager_total = None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    pairs_total = B * world * args.steps
    value = pairs_total / (ms_max / 1000.0)

    # ---------------- end to end through the public API with HOST buffers (`e2e`)
    # every step: pinned host inputs -> device, the step, results -> pinned host memory
    # (what evaluate.py:43-53 moves), all inside the timed region.
    if args.graph:
        from stitch_b200.pipeline import StreamedHotPath
        del out
        sp = StreamedHotPath(pb_host, size=SIZE, iters=ITERS, pyramid=True, device=dev)
        h2d, d2h = sp.h2d_bytes(), sp.d2h_bytes()

        def run_e2e(n):
            sp.fork(stream)
            for _ in range(n):
                res = sp.submit(pb_host)
            sp.join(stream)
            return res
        last_out = lambda: sp.slots[(sp.i - 1) % sp.depth]["out"]
    else:
        out_host = {k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
                    for k in ("warped_image_pred", "valid")}
        h2d = pb_host.nbytes()
        d2h = sum(h.numel() * h.element_size() for h in out_host.values())
        keep = {}

        def run_e2e(n):
            for _ in range(n):
                pbd = pb_host.map(lambda t: t.to(dev, non_blocking=True))
                o = hp.step(pbd)
                for k, h in out_host.items():
                    h.copy_(o[k], non_blocking=True)
                keep["out"] = o
            return out_host
        last_out = lambda: keep["out"]

    run_e2e(max(2, args.warmup // 2))
    barrier()
    w0 = time.perf_counter()
    e0.record(stream)
    host_result = run_e2e(args.steps)
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1000.0
    t = torch.tensor([max(e0.elapsed_time(e1), 0.0)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = pairs_total / (t.item() / 1000.0)
    out = last_out()
    # clocks were sampled from before the device-resident timed region to after the end-to-end one
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- the single collective of the path: final metric reduction
    # (computed from the results that arrived in pinned HOST memory)
    metric_sum = torch.tensor([host_result["warped_image_pred"].double().mean().item(), float(B)],
                              dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(metric_sum, op=dist.ReduceOp.SUM)

    if rank == 0:
        work = algorithmic_work(B, SIZE, ITERS, 256, pyramid=True)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # dominant kernel: corr_umma_kernel. Algorithmic bytes per launch = one direction of the batch:
        # B * (N1*N2*4 * (1 + 1/4 + 1/16 + 1/64) written + (N1+N2)*C*2 bf16 operands read)
        n = (SIZE // 8) ** 2
        gemm_bytes = B * (n * n * 4 * (1 + 0.25 + 0.0625 + 0.015625) + 2 * n * 256 * 2)
        gemm_flops = B * 2.0 * n * n * 256
        gemm_avg_ms = statistics.mean(gemm_ms) if gemm_ms else float("nan")
        achieved = gemm_bytes / (gemm_avg_ms / 1000.0) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "corr_umma_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, done = cpu_pairs_per_s(2, 12.0)
            cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "cpu_model": cpu_model(),
                   "sample": f"{done} pairs of the same op list (oracle port: numpy BLAS + OpenMP C), all host threads"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(world), "global_batch": B * world, "image_size": SIZE,
                       "lookup_iters": ITERS, "submission": ("cuda-graph replay" if args.graph else "eager launches") +
                       (", warp stage on a second stream (fork/join)" if args.overlap else ""),
                       "parallelism": f"pairs sharded x{world}, no data-path collective",
                       "l2": "per-step working set (2 x 1 GiB volumes) >> 126 MB L2, no explicit flush",
                       "algorithmic_bytes_per_step": work["bytes"], "algorithmic_flops_per_step": work["flops"]},
            "roofline": {"bound": "hbm", "kernel": "corr_umma_kernel<true> (tcgen05 cost volume + fused pyramid)",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": gemm_avg_ms,
                         "launches_timed": len(gemm_ms), "algorithmic_bytes_per_launch": gemm_bytes,
                         "tensor_tflops": gemm_flops / (gemm_avg_ms / 1000.0) / 1e12,
                         "tensor_frac_of_sustained": (gemm_flops / (gemm_avg_ms / 1000.0) / 1e12) /
                         float(peaks.get("bf16_tflops_sustained", 1400.0)),
                         "kernel_share_of_step": (2 * gemm_avg_ms / (ms_max / args.steps)) if gemm_ms else None,
                         "timed": ("CUDA events around 20 back-to-back launches on the launching stream right after the "
                                   "graph-replay region (events cannot be timed inside a captured graph)") if args.graph
                         else "CUDA events around every launch inside the timed region (eager)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "wall_ms_per_step": wall_ms / args.steps,
                    "how": ("double-buffered 3-stream pipeline (H2D | graph replay | D2H overlap)" if args.graph
                            else "serial H2D -> step -> D2H on one stream")},
            "gpu_launches": launches,
            "clocks": clocks,
            "reduced_metric": {"mean_warped_image": metric_sum[0].item() / world, "pairs": metric_sum[1].item()},
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="eager launches instead of replaying the step as one captured CUDA graph")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="run the warp stage after the cost-volume stage instead of on a second stream")
    ap.add_argument("--wc", action="store_true", help="write-combined pinned host input buffers (experiment)")
    ap.set_defaults(graph=True, overlap=True)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host thread
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = str(os.cpu_count())
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
