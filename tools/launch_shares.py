"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total
time and share of the captured window.
    python tools/launch_shares.py gpurun_out/launches.csv [--md]
"""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|unnamed>::|void |sf::", "", name)
    return name.split("(")[0][:90]


def main():
    path = sys.argv[1]
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        if r[ui] == "us":
            v *= 1e3
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / tot * 100:.1f}% |")
    print(f"| total | {len(rows) - 1} | {tot / 1e3:.1f} | 100% |")


if __name__ == "__main__":
    main()
