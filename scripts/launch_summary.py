#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`bench.py --frames 7 --steps 1 --warmup 1` into one table per kernel for the LAST GOP of the capture (markdown on stdout).
Usage: python scripts/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))[1:]
    byid = collections.OrderedDict()
    for r in rows:
        d = byid.setdefault(r[0], {"name": r[4].split("(")[0].replace("void ", "").replace("selfc::", ""), "grid": r[8]})
        d[r[12]] = float(r[14].replace(",", ""))
    L = list(byid.values())
    starts = [i for i, d in enumerate(L) if "fa_fwd" in d["name"]]
    G = L[starts[-1]:]
    agg = collections.OrderedDict()
    for d in G:
        a = agg.setdefault(d["name"][:48], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0) / 1e6
        a[2] += d.get("dram__bytes_read.sum", 0) / 1e9
        a[3] += d.get("dram__bytes_write.sum", 0) / 1e9
    tot = sum(a[1] for a in agg.values())
    print(f"{len(G)} launches in the last GOP of the capture, {tot:.3f} ms under ncu (cold caches, serialised launches)\n")
    print("| kernel | launches | ncu time (ms) | share | DRAM read (GB) | DRAM write (GB) | TB/s | us per launch |")
    print("|---|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[1] / tot:.3f} | {a[2]:.2f} | {a[3]:.2f} | {(a[2] + a[3]) / a[1]:.2f} | {a[1] / a[0] * 1000:.1f} |")


if __name__ == "__main__":
    main()
