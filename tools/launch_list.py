#!/usr/bin/env python
"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table kept under
profiles/.  usage: python tools/launch_list.py launches.csv out.md "<command the list was taken from>" """
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# Launch list of `{cmd}`\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` after the same command "
                "exited 0 without ncu.  Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n")
        f.write(f"{n} launches, {total:.1f} ms of device time; everything that is not `mk::` is the synthetic graph generator.\n\n")
        f.write("| kernel | launches | total ms | share | ms per launch |\n|---|---|---|---|---|\n")
        for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
            f.write(f"| `{name[:110]}` | {c} | {ms:.3f} | {100 * ms / total:.1f} % | {ms / c:.4f} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
