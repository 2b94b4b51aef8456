#!/usr/bin/env python
"""Opcode histogram per kernel of libvfmreg_b200.so (cuobjdump -sass), with the Blackwell-specific mnemonics called out:
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load, UTCBAR = tcgen05.commit, UTMAPF = bulk L2
prefetch, SYNCS = mbarrier, HMMA = mma.sync (legacy tensor path), DFMA = fp64.
    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vfm_registration_b200", "libvfmreg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "DFMA", "DMUL", "DADD",
       "MUFU", "FFMA", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "SHFL", "BAR", "ACQBULK", "ELECT")
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
        cur["_total"] += 1
        if m.group(1) in ("UTCHMMA", "UTMALDG", "UTCBAR") and m.group(2) and "2CTA" in m.group(2):
            cur[m.group(1) + ".2CTA"] += 1
print(f"# {os.path.basename(lib)}: {len(kernels)} kernels; instruction counts per kernel (static SASS), key mnemonics then the five most frequent")
for name, c in kernels.items():
    keys = " ".join(f"{k}={c[k]}" for k in KEY + ("UTCHMMA.2CTA", "UTMALDG.2CTA", "UTCBAR.2CTA") if c[k])
    top = " ".join(f"{k}:{v}" for k, v in c.most_common(7) if k != "_total" and k not in KEY)[:90]
    print(f"{name[:70]:70s} total={c['_total']:5d}  {keys}   | {top}")
