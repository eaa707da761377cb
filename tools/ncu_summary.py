#!/usr/bin/env python
"""Summarises ncu outputs from gpurun_out/ into small JSON files under profiles/ (run in the build container).

    python tools/ncu_summary.py <tag> --launches gpurun_out/launches.csv --reports gpurun_out/a.ncu-rep ...
"""
import argparse
import collections
import csv
import json
import subprocess

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
]


def launch_list(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(v[1] for v in agg.values())
    return {k: {"launches": v[0], "total_us": round(v[1], 1), "avg_us": round(v[1] / v[0], 2),
                "share": round(v[1] / total, 4)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}


def report(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")][:90]}
        for w in WANT:
            if w in hdr:
                d[w] = (row[hdr.index(w)] + " " + units[hdr.index(w)]).strip()
        out.append(d)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--reports", nargs="*", default=[])
    a = ap.parse_args()
    out = {}
    if a.launches:
        out["launch_list"] = launch_list(a.launches)
    for r in a.reports:
        out[r.split("/")[-1]] = report(r)
    path = f"profiles/{a.tag}_ncu_summary.json"
    json.dump(out, open(path, "w"), indent=1)
    print(path)
    if "launch_list" in out:
        for k, v in out["launch_list"].items():
            print(f"  {k:45s} n={v['launches']:4d} avg_us={v['avg_us']:8.2f} share={v['share']:.3f}")


if __name__ == "__main__":
    main()
