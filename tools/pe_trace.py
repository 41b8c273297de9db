"""PatchEmbed kernel pipeline trace: run a -DSB_PE_TRACE build (tools/pe_variants.sh "trace:-DSB_PE_TRACE") and print, for
iterations 16..23 of CTA 0, when each role started / finished (SM cycles relative to the conv1 issue of iteration 16)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tools", "probes", "libstitch_pe_trace.so")
os.environ["STITCH_B200_LIB"] = lib
sys.path.insert(0, ROOT)
import torch, stitch_b200 as sb
g = torch.Generator(device="cuda").manual_seed(7)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
maps = rnd(65536, 1, 64, 64) * 8
w1, b1, w2, b2, w3, b3 = rnd(16, 1, 6, 6) / 6, rnd(16) / 4, rnd(32, 16, 6, 6) / 24, rnd(32) / 4, rnd(64, 32, 6, 6) / 34, rnd(64) / 4
pack = sb.encoder.pack_patch_embed_weights(w1, w2, w3)
ms = []
for _ in range(int(os.environ.get("PE_TRACE_REPS", "3"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sb.encoder.patch_embed_proj(maps, w1, b1, w2, b2, w3, b3, pack=pack); e1.record()
    torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
h = ctypes.CDLL(lib)
buf = (ctypes.c_longlong * (128 + 296))()
assert h.sb_pe_trace_read(buf) == 0
names = ["c1 issue start", "c1 issue end", "c2 issue start", "c2 issue end", "c3 issue start", "c3 issue end",
         "E1a start", "E1a end", "E1b start", "E1b end", "E2 start", "E2 end", "E3 start", "E3 end", "ldr a1 free", "ldr done"]
t0 = buf[0]
print(os.path.basename(lib))
print("loop t  " + "".join("%16s" % n for n in names))
for t in range(8):
    print("t=%-5d " % (16 + t) + "".join("%16d" % (buf[t * 16 + e] - t0) for e in range(16)))
print("period (c1 issue start to start): " + " ".join(str(buf[(t + 1) * 16] - buf[t * 16]) for t in range(7)))
tot = [buf[128 + 2 * i] for i in range(148)]; loop = [buf[129 + 2 * i] for i in range(148)]
print("per-CTA kernel cycles: min %d median %d max %d;  loop only: min %d median %d max %d" % (
    min(tot), sorted(tot)[74], max(tot), min(loop), sorted(loop)[74], max(loop)))
print("loop cycles by CTA: " + " ".join(str(x // 1000) for x in loop))
print("last launch: %.1f us by CUDA events -> slowest CTA ran at %.0f MHz (clock64 cycles / event time)" % (ms[-1] * 1e3, max(tot) / ms[-1] / 1e3))
acc = (ctypes.c_longlong * (74 * 8))()
if hasattr(h, "sb_pe_acc_read") and h.sb_pe_acc_read(acc) == 0:
    print("per cluster (leader CTA): smid | loop kcycles | conv1 wait/issue | conv2 wait/issue | conv3 wait/issue (kcycles)")
    rows = sorted(range(74), key=lambda c: loop[2 * c])
    for c in rows:
        a = [acc[c * 8 + i] // 1000 for i in range(7)]
        print("cluster %2d smid %3d loop %5d | c1 %5d %5d | c2 %5d %5d | c3 %5d %5d" % (c, acc[c * 8 + 6], loop[2 * c] // 1000, a[0], a[1], a[2], a[3], a[4], a[5]))
