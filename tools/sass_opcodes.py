#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove (or disprove) a Blackwell-native kernel, from `cuobjdump -sass` of the in-tree
library: UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), and the legacy HMMA / LDSM (mma.sync / ldmatrix). Runs without a GPU.
    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "magpo_b200", "lib", "libmagpo_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA", "LDSM", "FFMA", "FFMA2"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts, total, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("magpo::", "")
        name = re.sub(r"^void ", "", name)
        cur = re.sub(r"\(.*", "", name)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for o in OPS:
            if op == o or (o not in ("FFMA", "HMMA") and op.startswith(o)) or (o == "HMMA" and op.startswith("HMMA")):
                counts[cur][o] += 1
print("SASS opcode counts per kernel of `magpo_b200/lib/libmagpo_b200.so` (`cuobjdump -sass`, sm_100a). tcgen05 = UTC*MMA + LDTM/STTM, "
      "TMA = UTMA*; HMMA/LDSM = legacy mma.sync / ldmatrix.\n")
print("| kernel | instructions | " + " | ".join(OPS) + " |")
print("|---|---|" + "---|" * len(OPS))
for k in sorted(total, key=lambda k: (-(counts[k]["UTCHMMA"] + counts[k]["UTCQMMA"]), -counts[k]["HMMA"], k)):
    if total[k] < 64 and not any(counts[k][o] for o in OPS[:11]):
        continue
    print(f"| `{k}` | {total[k]} | " + " | ".join(str(counts[k][o]) if counts[k][o] else "" for o in OPS) + " |")
