#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log: the launches of the LAST proof (from the last
r1cs_eval_kernel / or whole file), per launch with stream, and totals per kernel.  usage: launch_list.py file.csv [--all]"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, si, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Stream"), h.index("Grid Size")
    return [(r[ki].split("(")[0].replace("void ", ""), float(r[vi].replace(",", "")) / 1e3, r[si], r[gi]) for r in rows[1:]]


def main():
    seq = load(sys.argv[1])
    idx = [i for i, (k, *_) in enumerate(seq) if "r1cs_eval" in k]
    if idx and "--all" not in sys.argv:
        seq = seq[idx[-2]:idx[-1]] if len(idx) > 1 and "--prev" in sys.argv else seq[idx[-1]:]
    tot = collections.OrderedDict()
    for k, t, s, g in seq:
        if "--quiet" not in sys.argv:
            print(f"{s:>3} {t:9.1f} {g:>16} {k[:80]}")
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in tot.values())
    print(f"\n| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {n} | {t:.0f} | {100 * t / total:.1f}% |")
    print(f"\nTotal {total:.0f} us serialised, {len(seq)} launches.")


if __name__ == "__main__":
    main()
