"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel -> markdown table on stdout."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
        rows.append((name, ns))
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f} % | {t / c / 1e3:.1f} |")
    print(f"| **total** | {len(rows)} | {tot / 1e6:.3f} | 100 % | |")


if __name__ == "__main__":
    main(sys.argv[1])
