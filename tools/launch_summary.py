"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: the launches of the LAST evaluation
(after the last occurrence of the first kernel of an evaluation) or all launches grouped by kernel.
    python tools/launch_summary.py file.csv [first_kernel_substring]"""
import csv
import sys
from collections import OrderedDict


def rows(path):
    hdr = None
    for r in csv.reader(open(path, errors="replace")):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            yield dict(zip(hdr, r))


def short(name):
    name = name.replace("oo::<unnamed>::", "").replace("void oo::", "").replace("void ", "")
    return name.split("(")[0][:64]


def main():
    path = sys.argv[1]
    first = sys.argv[2] if len(sys.argv) > 2 else None
    rs = list(rows(path))
    if first:
        starts = [i for i, r in enumerate(rs) if first in r["Kernel Name"]]
        rs = rs[starts[-1]:] if starts else rs
    agg = OrderedDict()
    total = 0.0
    for r in rs:
        k = short(r["Kernel Name"]) + " " + r["Grid Size"].replace(" ", "")
        t = float(r["Metric Value"]) / 1e3
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    print(f"{len(rs)} launches, {total:.1f} us")
    for k, (n, t) in agg.items():
        print(f"{t:10.1f} us  x{n:<3d} {k}")


if __name__ == "__main__":
    main()
