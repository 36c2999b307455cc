"""Per-kernel SASS opcode counts of the shipped library (evidence that the hot kernels are tcgen05 / TMEM / TMA code):
    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.md
UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,
UBLKCP = cp.async.bulk (non-tensor), HMMA = mma.sync, LDGSTS = cp.async, LDSM = ldmatrix, MUFU = special-function unit."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "streamformer_b200", "lib", "libstreamformer_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "LDGSTS", "LDSM", "MUFU", "REDG", "ATOMS", "RED"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("sf::", "")
        name = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", name)        # drop the argument list
        kern = name
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[kern][o] += 1
print("# SASS opcode counts per kernel (cuobjdump -sass of streamformer_b200/lib/libstreamformer_b200.so, sm_100a)\n")
print("| kernel | instr | " + " | ".join(OPS) + " |")
print("|---|---|" + "---|" * len(OPS))
tot = collections.Counter()
for k, c in counts.items():
    if c["_total"] < 64 and not any(c[o] for o in OPS):
        continue
    print(f"| `{k[:110]}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    tot.update(c)
print(f"| **all {len(counts)} kernels** | {tot['_total']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")
