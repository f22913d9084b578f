#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list.

usage: python tools/summarize_ncu.py gpurun_out/launches.csv "<command that was profiled>" [--json out.json] [--skip-first N]

Prints per-kernel launch counts, total time and share (cold-cache, serialised: compare shares, not absolutes) and, when the
DRAM byte counters are present, measured DRAM traffic per launch.  --json writes, per convolution kernel family and tagged with the
digest of the build that was profiled, the numbers bench.py reports as roofline.traffic (entries of another build are ignored).
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name.replace("void ", "").strip()


def main():
    path, cmd = sys.argv[1], sys.argv[2]
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    launches = OrderedDict()       # id -> dict
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for row in csv.DictReader(lines):
        rec = launches.setdefault(row["ID"], {"kernel": short(row["Kernel Name"])})
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        metric = row["Metric Name"]
        if metric == "gpu__time_duration.sum":
            rec["us"] = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        elif metric.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            rec[metric] = val * mult
    agg = OrderedDict()
    for rec in launches.values():
        a = agg.setdefault(rec["kernel"], {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["launches"] += 1
        a["us"] += rec.get("us", 0.0)
        a["rd"] += rec.get("dram__bytes_read.sum", 0.0)
        a["wr"] += rec.get("dram__bytes_write.sum", 0.0)
    total = sum(a["us"] for a in agg.values()) or 1.0
    have_dram = any(a["rd"] + a["wr"] > 0 for a in agg.values())
    print("ncu launch list: %s" % cmd)
    print("(%d launches; cold-cache, serialised: compare shares, not absolutes)" % len(launches))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        line = "%-60s launches=%4d  total_us=%10.1f  share=%5.1f%%" % (k[:60], a["launches"], a["us"], 100 * a["us"] / total)
        if have_dram:
            line += "  dram_MB/launch rd=%8.2f wr=%8.2f" % (a["rd"] / a["launches"] / 1e6, a["wr"] / a["launches"] / 1e6)
        print(line)
    fams = OrderedDict()
    for k, a in agg.items():
        # family names as bench.py reports them (roofline.kernel / roofline.families)
        name = next((fam for fam in ("conv_tc_kernel", "conv_halo_kernel", "conv1x1_kernel", "conv_dual_kernel", "stem_tc_kernel") if fam in k), None)
        if name is None:
            continue
        fam = fams.setdefault(name, {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        for key in fam:
            fam[key] += a[key]
    for name, fam in fams.items():
        print("%s family: launches=%d share=%.1f%%" % (name, fam["launches"], 100 * fam["us"] / total)
              + ("  measured DRAM bytes/launch = %.1f MB (read %.1f + write %.1f)" % ((fam["rd"] + fam["wr"]) / fam["launches"] / 1e6,
                 fam["rd"] / fam["launches"] / 1e6, fam["wr"] / fam["launches"] / 1e6) if have_dram else ""))
    if out_json and fams and have_dram:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from hawq_b200.build import OUT
        try:
            build = open(OUT + ".stamp").read().strip()[:16]
        except OSError:
            build = None
        with open(out_json, "w") as f:
            json.dump({"command": cmd, "build": build,
                       "kernels": {name: {"launches": fam["launches"], "share_of_listed_time": fam["us"] / total,
                                          "dram_bytes_read_per_launch": fam["rd"] / fam["launches"],
                                          "dram_bytes_write_per_launch": fam["wr"] / fam["launches"],
                                          "traffic_bytes_per_launch": (fam["rd"] + fam["wr"]) / fam["launches"]} for name, fam in fams.items()}}, f, indent=1)


if __name__ == "__main__":
    main()
