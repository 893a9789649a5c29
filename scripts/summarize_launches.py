#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: calls, total us, share."""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    m = re.match(r"([A-Za-z0-9_:]+)(<[^(]*>)?", name)
    base = m.group(1) if m else name
    tmpl = (m.group(2) or "") if m else ""
    tmpl = tmpl.replace("__nv_bfloat16", "bf16")
    return base + (tmpl if len(tmpl) < 40 else "")


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        k = short(row["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised under ncu)")
    print(f"{'kernel':60s} {'calls':>6s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} {n:6d} {t:10.1f} {t / n:8.1f} {100 * t / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
