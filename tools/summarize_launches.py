#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.

  python tools/summarize_launches.py gpurun_out/launches.csv [--step-marker KERNEL_SUBSTR] > profiles/rN_..._summary.md

With --step-marker the list is cut into steps at every launch whose name contains the marker
(default: the first kernel of a bench step, `reduce_kernel<512, 1, native::ReduceOp<float, native::MinOps`)
and the LAST complete step is summarised; otherwise the whole list is.
"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    marker = None
    if "--step-marker" in sys.argv:
        marker = sys.argv[sys.argv.index("--step-marker") + 1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        rows.append((r[ki], ms))
    if marker:
        cuts = [i for i, (k, _) in enumerate(rows) if marker in k]
        if "--step-index" in sys.argv:                      # the step that starts at the given occurrence of the marker
            j = int(sys.argv[sys.argv.index("--step-index") + 1])
            rows = rows[cuts[j]:cuts[j + 1]]
        elif len(cuts) >= 2:
            rows = rows[cuts[-2]:cuts[-1]]
    agg = OrderedDict()
    for k, ms in rows:
        a = agg.setdefault(k, [0.0, 0])
        a[0] += ms
        a[1] += 1
    total = sum(a[0] for a in agg.values())
    print(f"{len(rows)} launches, {total:.3f} ms of kernel time\n")
    print("| ms | share | launches | kernel |\n|---|---|---|---|")
    for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
        print(f"| {ms:.3f} | {100 * ms / total:.1f}% | {n} | `{k[:110]}` |")


if __name__ == "__main__":
    main()
