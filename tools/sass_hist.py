#!/usr/bin/env python
"""SASS opcode histogram of libfpmb200.so per kernel family (cuobjdump -sass), the evidence for the TMA / DSMEM /
packed-fp32 claims in DESIGN.md: UTMALDG/UTMASTG (cp.async.bulk.tensor), STAS (st.async to a peer CTA's shared
memory), SYNCS (mbarrier), FFMA2/FMUL2/FADD2 (packed fp32x2), LDS/STS, REDUX, MUFU.
    python tools/sass_hist.py [lib] > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fpm-opencv_b200", "lib", "libfpmb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEYS = ["UTMALDG", "UTMASTG", "UTMAPF", "STAS", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "LDS", "STS",
        "LDG", "STG", "REDUX", "SHFL", "BAR", "ATOMS", "IMAD", "LOP3", "UCGABAR", "MEMBAR", "HMMA", "UTCMMA"]
per = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op == k or (k in ("UTMALDG", "UTMASTG", "SYNCS", "LDS", "STS", "LDG", "STG", "MUFU", "ATOMS", "BAR") and op.startswith(k)):
                per[cur][k] += 1
                break
print("# %s  (sm_100a SASS, cuobjdump)" % os.path.relpath(lib, ROOT))
tot = collections.Counter()
for fn, c in per.items():
    name = demangle(fn)
    name = re.sub(r"\(.*", "", name)
    print("%-120s total %6d | %s" % (name[:120], c["_total"], " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))
    tot.update(c)
print("%-120s total %6d | %s" % ("ALL KERNELS", tot["_total"], " ".join("%s=%d" % (k, tot[k]) for k in KEYS if tot[k])))
