"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^.*::", "", name)
        rows.append((name, val * scale))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    print("# %s: %d launches, %.1f us total (ncu per-launch times are cold-cache and serialised: compare shares)" % (path, len(rows), tot))
    print("%-60s %8s %12s %10s %8s" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %8d %12.1f %10.2f %7.1f%%" % (n[:60], c, v, v / c, 100 * v / tot))


if __name__ == "__main__":
    main(sys.argv[1])
