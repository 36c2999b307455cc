"""Top stalled SASS lines of one launch in an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_hot.py rep.ncu-rep <launch-index> [top-n]
"""
import csv
import io
import subprocess
import sys

rep, launch = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
h = rows[hdr_i]


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


data, seen = [], set()
for r in rows[hdr_i + 1:]:
    if len(r) != len(h) or r[0] == "Address" or r[0] in seen:
        continue
    seen.add(r[0])
    data.append(r)
key, src = h.index("# Samples"), h.index("Source")
st = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(f(r[key]) for r in data)
print(f"launch {launch}: {int(tot)} samples, {len(data)} SASS lines")
agg = {}
for r in data:
    for i in st:
        agg[h[i]] = agg.get(h[i], 0) + f(r[i])
print("stall mix:", ", ".join(f"{k[6:]} {v / tot * 100:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in enumerate(sorted(data, key=lambda r: -f(r[key]))[:topn]):
    stalls = sorted([(f(r[i]), h[i][6:]) for i in st], reverse=True)[:2]
    idx = data.index(r)
    print(f"{f(r[key]) / tot * 100:5.1f}%  #{idx:4d} {r[src][:86]:86s} {stalls[0][1]}:{int(stalls[0][0])} {stalls[1][1]}:{int(stalls[1][0])}")
