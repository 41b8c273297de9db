"""profiles/sass_summary.txt: per-kernel counts of the SASS mnemonics that prove what the shipped library runs on —
tcgen05 MMAs (UTCHMMA), tensor-memory loads / copies (LDTM, UTCCP), TMA loads / stores (UTMALDG, UTMASTG, UBLKCP),
mbarriers (SYNCS), packed fp32 (FFMA2 ...).  Runs where the library is built (no GPU needed):
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "seamless-through-breaking-rethinking-image-stitching-for-optimal-alignment_b200", "lib", "libstitchb200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTCCP", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "ELECT", "FFMA2", "FMUL2", "FADD2",
        "MUFU", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "SHFL", "REDUX"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["_total"] += 1
            base = op.split(".")[0]
            per[cur][base] += 1
            if op.startswith("UTCHMMA.2CTA"):
                per[cur]["UTCHMMA.2CTA"] += 1
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print("# " + "  ".join(KEYS))
    tot = collections.Counter()
    for name, cnt in per.items():
        cols = [f"{k}={cnt[k]}" for k in KEYS if cnt[k]]
        print(f"{name}\n    instructions={cnt['_total']}  " + "  ".join(cols))
        tot.update(cnt)
    print("\n# whole library: " + "  ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    main()
