"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total time / share,
and (with --seq) the launch sequence between two adam_kernel launches (= one train step).
usage: python tools/launch_summary.py launches.csv [--seq] [--step]"""
import collections
import csv
import re
import sys


def short(n):
    n = re.sub(r"\(.*", "", n)
    return n.replace("void ", "")[:70]


def main():
    rows = []
    for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')):
        if r[0] == "ID":
            continue
        rows.append((short(r[4]), float(r[-1]) / 1e3, r[8]))
    if "--step" in sys.argv:
        idx = [i for i, r in enumerate(rows) if "adam_kernel" in r[0]]
        # one step = after the first complete adam group .. through the next one
        starts = [i for k, i in enumerate(idx) if k == 0 or idx[k - 1] != i - 1]
        if len(starts) >= 2:
            rows = rows[starts[0] + 2:starts[1] + 2]
    if "--seq" in sys.argv:
        for i, r in enumerate(rows):
            print(f"{i:4d} {r[1]:9.1f} us  {r[2]:>14s}  {r[0]}")
        return
    agg = collections.OrderedDict()
    for n, t, _ in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print(f"# {len(rows)} launches, {tot / 1e3:.3f} ms summed kernel time")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:72s} n={c:4d} {t:10.1f} us {100 * t / tot:5.1f}%")


main()
