#!/usr/bin/env python
"""Per-kernel SASS digest of libsplitvae.so (cuobjdump -sass): instruction count and the mnemonics that prove which hardware paths a
kernel uses - UTCHMMA (tcgen05.mma kind::f16), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA load / store), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), plus registers from -res-usage.  Runs on the build host (no GPU needed):

    python scripts/sass_digest.py > profiles/r02_sass_digest.txt
"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "splitvae_b200", "libsplitvae.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "BAR"]


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
        return [o.strip() for o in out[:len(names)]]
    except OSError:
        return names


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    kernels = OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), {"n": 0, **{k: 0 for k in KEYS}})
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur["n"] += 1
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    cur[k] += 1
    names = list(kernels)
    pretty = demangle(names)
    print(f"# SASS digest of {os.path.relpath(LIB, ROOT)} (sm_100a), {len(names)} kernels; counts are static instructions")
    print(f"{'kernel':58s} {'instr':>6s} {'regs':>5s} " + " ".join(f"{k:>7s}" for k in KEYS))
    tot = {k: 0 for k in KEYS}
    for n, p in zip(names, pretty):
        k = kernels[n]
        short = re.sub(r"\(.*", "", p).replace("sv::(anonymous namespace)::", "").replace("sv::", "")[:58]
        print(f"{short:58s} {k['n']:6d} {regs.get(n, (0, 0))[0]:5d} " + " ".join(f"{k[x]:7d}" for x in KEYS))
        for x in KEYS:
            tot[x] += k[x]
    print(f"{'TOTAL':58s} {sum(k['n'] for k in kernels.values()):6d} {'':5s} " + " ".join(f"{tot[x]:7d}" for x in KEYS))


if __name__ == "__main__":
    main()
