#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum per launch, --csv) per kernel.
Usage: python scripts/summarize_launches.py gpurun_out/launches.csv "<command that was profiled>" """
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else ""
    print("# ncu launch list: ncu --metrics gpu__time_duration.sum --clock-control none --csv " + cmd)
    print("# cold-cache, serialised per-launch times: compare SHARES with bench.py's live CUDA-event numbers, not absolutes")
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        name = r[hdr.index("Kernel Name")].split("(")[0][:60]
        val = float(r[hdr.index("Metric Value")].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        val = val / 1e3 if unit == "ns" else val * 1e3 if unit == "ms" else val
        agg[name][0] += 1
        agg[name][1] += val
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s launches=%4d total_us=%10.1f share=%5.1f%% avg_us=%8.1f" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


if __name__ == "__main__":
    main()
