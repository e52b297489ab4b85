#!/usr/bin/env python
"""Per-kernel SASS instruction histogram of the product library (cuobjdump -sass; no GPU needed): the mnemonics that prove
what a kernel is made of -- UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA loads, .IM2COL for im2col maps),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), LDGSTS (cp.async), HMMA/FFMA for contrast -- plus the top mnemonics overall.
Usage: sass_histogram.py [lib.so] > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "minerva_b200", "lib", "libmnv_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "HMMA", "IMMA", "FFMA", "DFMA",
       "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]
kern, hist, full = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("mnv::", "")
        hist[kern] = collections.Counter()
        full[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        op, mods = m.group(1), m.group(2)
        full[kern][op] += 1
        if op in KEY:
            tag = op
            if op == "UTMALDG":
                tag += "".join(x for x in (".2D", ".3D", ".4D", ".IM2COL") if x in mods)
            if op == "UTCHMMA" and ".2CTA" in mods:
                tag += ".2CTA"
            hist[kern][tag] += 1
print("SASS instruction histogram of %s (sm_100a), %d kernels" % (os.path.basename(lib), len(hist)))
print("tensor-core / TMA kernels:")
for k in sorted(hist, key=lambda k: -sum(full[k].values())):
    h = hist[k]
    if any(t.startswith(("UTCHMMA", "UTMALDG", "LDTM")) for t in h):
        print("  %-60s %6d instr  %s" % (k[:60], sum(full[k].values()), "  ".join("%s=%d" % kv for kv in sorted(h.items()))))
print("other kernels (no tensor-core / TMA instruction):")
for k in sorted(hist, key=lambda k: -sum(full[k].values())):
    h = hist[k]
    if not any(t.startswith(("UTCHMMA", "UTMALDG", "LDTM")) for t in h):
        print("  %-60s %6d instr  %s" % (k[:60], sum(full[k].values()), "  ".join("%s=%d" % kv for kv in sorted(h.items()) if kv[0] in ("LDG", "STG", "LDS", "STS", "SHFL", "LDGSTS", "FFMA", "DFMA", "MUFU", "BAR"))))
tot = collections.Counter()
for k in full:
    tot.update(full[k])
print("library-wide: " + "  ".join("%s=%d" % kv for kv in tot.most_common(24)))
forbidden = [k for k in hist if any(x in k.lower() for x in ("cublas", "cudnn", "cutlass", "triton"))]
print("library kernels from cuBLAS / cuDNN / CUTLASS / Triton: %s" % (forbidden or "none"))
