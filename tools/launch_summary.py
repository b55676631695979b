#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if r]
    hi = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[hi]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        name = r[ik]
        val = float(r[iv].replace(",", ""))
        unit = r[iu]
        ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9}.get(unit, 1)
        short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        if short.startswith("at::") or "at::native" in short or "elementwise" in short:
            short = "[torch: synthetic-data generation / copies] " + short.split("<")[0]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot_ours = sum(v[1] for k, v in agg.items() if not k.startswith("[torch"))
    print("| kernel | launches | total ms | mean ms | share of apgpu kernel time |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = f"{100 * ns / tot_ours:.1f} %" if not k.startswith("[torch") else "-"
        print(f"| `{k[:110]}` | {n} | {ns / 1e6:.3f} | {ns / 1e6 / n:.3f} | {share} |")


if __name__ == "__main__":
    main()
