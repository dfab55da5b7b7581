"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':74s} {'n':>5s} {'avg_us':>9s} {'total_us':>10s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:74s} {len(v):5d} {sum(v) / len(v):9.2f} {sum(v):10.1f} {sum(v) / tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
