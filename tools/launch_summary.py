#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv --log-file X.csv`): launches, total / mean duration, share of the profiled region, DRAM MB per launch.

  python tools/launch_summary.py gpurun_out/launches.csv "header comment" > profiles/rNN_launches_summary.csv
"""
import collections
import csv
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6,
        "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    hdr, rows = None, []
    for r in csv.reader(open(path)):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) == len(hdr):
            rows.append(r)
    ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = collections.OrderedDict()
    launch = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
        val = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        d = launch.setdefault(r[ii], {"name": name})
        d[r[mi]] = val
    for d in launch.values():
        a = per.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in per.values())
    if note:
        print(f"# {note}")
    print(f"# total {total:.0f} us over {sum(a[0] for a in per.values())} launches")
    print("launches,total_us,us_per_launch,share_pct,dram_read_MB_per_launch,dram_write_MB_per_launch,kernel")
    for name, a in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[0]},{a[1]:.1f},{a[1] / a[0]:.1f},{100 * a[1] / total:.1f},{a[2] / a[0]:.1f},{a[3] / a[0]:.1f},{name}")


if __name__ == "__main__":
    main()
