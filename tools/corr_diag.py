"""GPU bring-up diagnostics for the tcgen05 cost-volume kernel: runs a ladder of
shapes from one tile / one K panel upwards and, on a mismatch, prints enough
structure (per-block match maps, permutation probes) to localise a descriptor
or swizzle bug from one run."""
import sys, os, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200
from stitch_b200 import _lib, corr as C


def ref_bf16(f1, f2):
    b, c = f1.shape[:2]
    a = f1.bfloat16().float().reshape(b, c, -1)
    d = f2.bfloat16().float().reshape(b, c, -1)
    return torch.bmm(a.transpose(1, 2).double(), d.double()).float()


def blockmap(ok, bs=32):
    h, w = ok.shape
    rows = []
    for y in range(0, min(h, 256), bs):
        rows.append(" ".join(f"{ok[y:y+bs, x:x+bs].mean():4.2f}" for x in range(0, min(w, 256), bs)))
    return "\n".join(rows)


def run(b, c, h1, w1, h2, w2, lv=0):
    g = torch.Generator(device="cuda").manual_seed(b * 1000 + c + h1 + w2)
    f1 = torch.randn(b, c, h1, w1, device="cuda", generator=g)
    f2 = torch.randn(b, c, h2, w2, device="cuda", generator=g)
    t0 = time.time()
    res = C.corr(f1, f2, pyramid_levels=lv)
    vol = res[0] if lv else res
    torch.cuda.synchronize()
    dt = time.time() - t0
    v = vol.reshape(b, h1 * w1, h2 * w2)
    r = ref_bf16(f1, f2)
    err = (v - r).abs().max().item()
    scale = r.abs().max().item()
    status = "OK " if err <= 1e-4 * scale else "BAD"
    print(f"[{status}] B={b} C={c} N1={h1*w1} N2={h2*w2} lv={lv}: max err {err:.3e} (scale {scale:.1f}) {dt*1e3:.1f} ms dbg={_lib.load().sb_debug_word():#x}", flush=True)
    if status == "BAD":
        ok = ((v[0] - r[0]).abs() <= 1e-3 * scale).cpu().numpy()
        print("match fraction per 32x32 block (first 256x256 of batch 0):")
        print(blockmap(ok))
        vv, rr = v[0].cpu().numpy(), r[0].cpu().numpy()
        print("row 0, first 16 got :", np.round(vv[0, :16], 2))
        print("row 0, first 16 want:", np.round(rr[0, :16], 2))
        print("row 1, first 16 got :", np.round(vv[1, :16], 2))
        print("row 1, first 16 want:", np.round(rr[1, :16], 2))
        # is the output a permutation of the reference inside 128-byte groups (store swizzle bug)?
        n = min(vv.shape[1], 32)
        for chunk_xor in range(8):
            perm = np.arange(n).reshape(-1, 4)
            idx = (np.arange(n // 4) ^ chunk_xor)
            cand = rr[:8, :n].reshape(8, -1, 4)[:, idx % (n // 4), :].reshape(8, n)
            print(f"  chunk-xor {chunk_xor}: rows0-7 match {np.mean(np.abs(cand - vv[:8, :n]) < 1e-2 * scale):.2f}")
        # transposed?
        if vv.shape[0] == vv.shape[1]:
            print("  matches reference^T:", float(np.mean(np.abs(vv - rr.T) < 1e-2 * scale)))
        # zero / nan stats
        print("  zeros:", float((vv == 0).mean()), "nans:", float(np.isnan(vv).mean()))
    if lv:
        cm = v.reshape(-1, 1, h2, w2)
        want = cm
        for l in range(lv):
            want = torch.nn.functional.avg_pool2d(want, 2, stride=2)
            e = (res[1][l] - want).abs().max().item()
            print(f"      pyramid level {l+1}: max err vs avg_pool2d(own volume) {e:.3e}")
    return status == "OK "


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "device_check", _lib.load().sb_device_check())
    ladder = [(1, 64, 8, 16, 8, 16), (1, 64, 16, 16, 16, 16), (1, 128, 8, 16, 8, 16), (1, 256, 8, 16, 8, 16),
              (1, 256, 16, 16, 32, 16), (2, 256, 16, 32, 32, 32), (1, 100, 5, 7, 6, 6), (1, 256, 64, 64, 64, 64),
              (16, 256, 64, 64, 64, 64)]
    allok = True
    for s in ladder:
        allok &= run(*s)
    allok &= run(1, 256, 64, 64, 64, 64, lv=3)
    allok &= run(16, 256, 64, 64, 64, 64, lv=3)
    # timing of the big case
    g = torch.Generator(device="cuda").manual_seed(0)
    f1 = torch.randn(16, 256, 64, 64, device="cuda", generator=g)
    f2 = torch.randn(16, 256, 64, 64, device="cuda", generator=g)
    t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
    for lv in (0, 3):
        for _ in range(3):
            C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=lv)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=lv)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byt = 16 * (4096 * 4096 * 4 * (1 + (0.328125 if lv else 0)) + 2 * 4096 * 256 * 2)
        print(f"corr_tokens B=16 lv={lv}: {ms:.3f} ms/launch, {byt/ms/1e6:.0f} GB/s algorithmic, {16*2*4096*4096*256/ms/1e9:.0f} TFLOP/s")
    e0.record()
    for _ in range(10):
        C.tokens_bf16(f1)
    e1.record(); torch.cuda.synchronize()
    print(f"tokens_bf16 B=16: {e0.elapsed_time(e1)/10:.3f} ms")
    print("ALL OK" if allok else "SOME BAD")
    sys.exit(0 if allok else 1)
