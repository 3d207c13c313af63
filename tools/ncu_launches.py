#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file x.csv`).

  python tools/ncu_launches.py gpurun_out/x_launches.csv "comment" ... > profiles/rNN_ncu_launches_....txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    for c in sys.argv[2:]:
        print("# " + c)
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    head = rows[0]
    ki, mi, vi, ui = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value"), head.index("Metric Unit")
    tot = OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = re.sub(r"\(.*", "", r[ki])
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1; t[1] += v; n += 1
    all_us = sum(t[1] for t in tot.values())
    print(f"# launches captured: {n}, total {all_us / 1e3:.3f} ms (cold-cache, serialised: compare shares, not absolutes)")
    for name, (k, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:<70} launches {k:5d}  total {us / 1e3:10.3f} ms  share {100 * us / all_us:6.2f}%  avg {us / k:10.2f} us")


if __name__ == "__main__":
    main()
