#!/usr/bin/env python3
"""Opcode histogram of every kernel of libav1b200.so (cuobjdump -sass), as Markdown on stdout:
memory-instruction widths (LDG/STG/LDS/STS .128/.64/32/.U8), TMA / bulk-copy (UTMALDG, UBLKCP),
packed integer (IDP, VIADD, VIMNMX, VABSDIFF4, PRMT), barriers.  Usage: tools/sass_histogram.py [lib]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "av1dec_b200", "lib", "libav1b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
groups = [("total", lambda o: True), ("LDG.128", lambda o: o.startswith("LDG") and ".128" in o), ("LDG.64", lambda o: o.startswith("LDG") and ".64" in o),
          ("LDG.U8/S8", lambda o: o.startswith("LDG") and (".U8" in o or ".S8" in o)), ("LDG other", lambda o: o.startswith("LDG") and not any(t in o for t in (".128", ".64", ".U8", ".S8"))),
          ("STG.128", lambda o: o.startswith("STG") and ".128" in o), ("STG.64", lambda o: o.startswith("STG") and ".64" in o), ("STG.U8", lambda o: o.startswith("STG") and ".U8" in o),
          ("STG other", lambda o: o.startswith("STG") and not any(t in o for t in (".128", ".64", ".U8"))), ("LDS", lambda o: o.startswith("LDS")), ("STS", lambda o: o.startswith("STS")),
          ("UTMALDG", lambda o: o.startswith("UTMALDG")), ("UBLKCP", lambda o: o.startswith("UBLKCP")), ("SYNCS", lambda o: o.startswith("SYNCS")),
          ("IDP", lambda o: o.startswith("IDP")), ("VIADD/VIMNMX/VABSDIFF", lambda o: o.startswith(("VIADD", "VIMNMX", "VABSDIFF", "VIADDMNMX"))),
          ("PRMT", lambda o: o.startswith("PRMT")), ("IMAD", lambda o: o.startswith("IMAD")), ("BAR", lambda o: o.startswith("BAR")), ("SHFL/REDUX", lambda o: o.startswith(("SHFL", "REDUX")))]
print("| kernel | " + " | ".join(g for g, _ in groups) + " |")
print("|---|" + "---|" * len(groups))
for k, c in kernels.items():
    print("| `%s` | " % k[:60] + " | ".join(str(sum(n for o, n in c.items() if f(o))) for _, f in groups) + " |")
