"""Per CUDA source line stall samples of one launch in an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py rep.ncu-rep <launch-index> [top-n]
"""
import csv
import io
import subprocess
import sys

rep, launch = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", str(launch),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
samp = h.index("# Samples")
stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
lines = []
fpath = ""
for r in rows[:hi]:
    if r and r[0] == "File Path":
        fpath = r[1]
for r in rows[hi + 1:]:
    if len(r) != len(h):
        if r and r[0] == "File Path":
            fpath = r[1]
        continue
    if r[0] == "Line No" or r[0] == "":
        continue
    try:
        n = float(r[samp])
    except ValueError:
        continue
    top = sorted(((float(r[i] or 0), h[i][6:]) for i in stalls), reverse=True)[:2]
    lines.append((n, fpath.split("/")[-1], r[0], r[1].strip()[:100], top))
tot = sum(l[0] for l in lines)
print(f"launch {launch}: {int(tot)} samples over {len(lines)} source lines")
for n, f, ln, src, top in sorted(lines, key=lambda l: -l[0])[:topn]:
    print(f"{n / tot * 100:5.1f}%  {f}:{ln:>4s}  {src:100s} {top[0][1]}:{int(top[0][0])} {top[1][1]}:{int(top[1][0])}")
