#!/usr/bin/env python3
"""profiles/sass_summary.txt: per kernel of orb_slam_b200/liborbfe.so, the SASS mnemonics that show what kind of code it is
(cuobjdump -sass, static counts): TMA (UTMALDG / UBLKCP), mbarrier (SYNCS), packed integer min/max (VIMNMX3 / VIMNMX),
integer dot products (IDP), population count (POPC), LOP3, peer-visible release stores (ST.E.STRONG.SYS / MEMBAR.SYS),
tensor-core mnemonics (HMMA / UTC*MMA: must be zero, there is no dense contraction on this path)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "orb_slam_b200", "liborbfe.so")
WATCH = ["UTMALDG", "UBLKCP", "SYNCS", "VIMNMX3", "VIMNMX", "HFMA2", "HADD2", "IDP", "POPC", "LOP3", "PRMT", "IMAD", "SHFL", "LDS", "LDG", "STG", "ATOMS", "ATOMG",
         "RED", "MEMBAR", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "DFMA", "DADD", "DMUL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op, mods = m.group(1), m.group(2)
            kernels[cur][op] += 1
            kernels[cur]["_total"] += 1
            if op in ("ST", "STG") and ".SYS" in mods:
                kernels[cur]["ST.SYS"] += 1
            if op == "MEMBAR" and ".SYS" in mods:
                kernels[cur]["MEMBAR.SYS"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    arch = re.findall(r"arch = (sm_\w+)", txt)
    print("liborbfe.so: %d kernels, arch %s" % (len(kernels), sorted(set(arch))))
    print("%-64s %7s  %s" % ("kernel", "instrs", "watched mnemonics (static counts)"))
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\((?!anonymous).*", "", dn).replace("orbfe::", "").replace("(anonymous namespace)::", "")
        items = ["%s=%d" % (k, c[k]) for k in WATCH + ["ST.SYS", "MEMBAR.SYS"] if c.get(k)]
        print("%-64s %7d  %s" % (short[:64], c["_total"], " ".join(items)))
    tens = sum(c.get(k, 0) for c in kernels.values() for k in ("HMMA", "UTCHMMA", "UTCQMMA", "LDTM"))
    print("tensor-core mnemonics in the library: %d (expected 0: no dense contraction on this path)" % tens)


if __name__ == "__main__":
    main()
