#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv --log-file X.csv`): launches, total / mean duration, share of the profiled region, DRAM MB per launch.

  python tools/launch_summary.py gpurun_out/launches.csv "header comment" > profiles/rNN_launches_summary.csv
  python tools/launch_summary.py --traffic gpurun_out/launches.csv > profiles/rNN_gemm_traffic.json
      (DRAM bytes per launch of the DiT GEMMs of the denoise step: every launch before the first VAE kernel)
"""
import collections
import csv
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6,
        "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def traffic(path):
    import json
    hdr, launch = None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
                ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
            continue
        if len(r) != len(hdr):
            continue
        d = launch.setdefault(r[ii], {"name": re.sub(r"\(.*", "", r[ki]).split("::")[-1]})
        d[r[mi]] = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
    step = []
    for d in launch.values():  # the denoise step ends where the decode starts
        if d["name"].startswith(("denorm_kernel", "vae_")):
            break
        step.append(d)
    gemm = [d for d in step if d["name"].startswith(("gemm_bf16_tn_kernel", "gemm_pair_bf16_tn_kernel"))]
    per = collections.OrderedDict()
    for d in gemm:
        a = per.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["gpu__time_duration.sum"]
        a[2] += d["dram__bytes_read.sum"] * 1e6
        a[3] += d["dram__bytes_write.sum"] * 1e6
    step_us = sum(d["gpu__time_duration.sum"] for d in step)
    out = {
        "source": f"{path} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                  "over tools/profile_step2.py: ONE batched-CFG denoise step at c2 = the forward bench.py times, M = 9984 rows "
                  "per GEMM; ncu flushes the caches before every launch, so the bytes are cold-cache upper bounds)",
        "dit_step_launches": len(step),
        "dit_step_us_under_ncu": step_us,
        "gemm_launches": len(gemm),
        "gemm_share_of_step": sum(a[1] for a in per.values()) / step_us,
        "bytes_per_launch_mean": sum(a[2] + a[3] for a in per.values()) / max(len(gemm), 1),
        "per_kernel": {k: {"launches": a[0], "us_per_launch": a[1] / a[0], "dram_read_bytes_per_launch": a[2] / a[0],
                           "dram_write_bytes_per_launch": a[3] / a[0]} for k, a in per.items()},
    }
    print(json.dumps(out, indent=1))


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[2])
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
